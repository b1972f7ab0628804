source tools/gpu_ab.sh
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
run base A=1
BENCH_ARGS="--workload 1024x128x1024" run slabshape A=1
