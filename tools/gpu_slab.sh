#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): slab parity check, then the bench at N ranks.
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/slab_check.py > gpurun_out/slab_check_$N.log 2>&1; echo "slab_check exit $?" >> gpurun_out/slab_check_$N.log
grep -E "FAIL|PASSED|FAILED|Error|error" gpurun_out/slab_check_$N.log | head -20; tail -5 gpurun_out/slab_check_$N.log
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_slab_$N.json 2> gpurun_out/bench_slab_$N.err
cat gpurun_out/bench_slab_$N.json; tail -5 gpurun_out/bench_slab_$N.err
