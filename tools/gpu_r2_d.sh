#!/bin/bash
# Round 2, GPU pass D: fused Poisson path + fp32 mom3 + surface model (short timeouts: a hang must not burn the budget).
set -x
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fft_roundtrip" > gpurun_out/pytest_fft.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_fft.log
tail -5 gpurun_out/pytest_fft.log
if [ $rc -eq 124 ]; then echo "HANG in the fused Poisson path: stopping"; exit 1; fi
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_surface.py -m gpu -q -x -k "pres_2 or full_rk3_step or surface or self_driven" > gpurun_out/pytest_pres.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_pres.log
tail -12 gpurun_out/pytest_pres.log
if [ $rc -eq 124 ]; then echo "HANG: stopping"; exit 1; fi
run() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-side-configs --workload 512x512x512 ${BENCH_ARGS} > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/ab_{tag}.json'))
    print(tag, f"{d['ms_per_step']:.2f} ms/step", d['clocks'], d.get('post_step_divergence'), ' '.join(f"{k.replace('_kernel','')}={v:.2f}" for k,v in d['kernels_ms_per_step'].items()))
except Exception as e:
    print(tag, 'FAILED', e, open(f'gpurun_out/ab_{tag}.err').read()[-700:])
PY
}
run fused
run unfused MHH_PRES_FUSED=0
BENCH_ARGS="--dtype f32" run f32_fused
BENCH_ARGS="--dtype f32 --igc 3" run f32_igc3
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
timeout 200 python tools/fft_compare.py > gpurun_out/fft_compare_f64.json 2> gpurun_out/fft_compare.err; cat gpurun_out/fft_compare_f64.json
timeout 200 python tools/fft_compare.py --dtype f32 > gpurun_out/fft_compare_f32.json 2>> gpurun_out/fft_compare.err
timeout 300 python tools/ref_cuda_bench.py > gpurun_out/ref_cuda_bench_f64.json 2> gpurun_out/ref_cuda_bench.err; cat gpurun_out/ref_cuda_bench_f64.json; tail -3 gpurun_out/ref_cuda_bench.err
du -sh gpurun_out
