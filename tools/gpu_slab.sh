#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): slab parity check (fused peer transposes, then NCCL all-to-all), then the bench in both modes.
set -x
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/slab_check.py > gpurun_out/slab_check_$N.log 2>&1; echo "slab_check (peer) exit $?" >> gpurun_out/slab_check_$N.log
grep -E "FAIL|PASSED|FAILED|Error|error|exit" gpurun_out/slab_check_$N.log | head -20
if [ "${SKIP_NCCL_CHECK:-0}" != "1" ]; then
MHH_NO_PEER=1 timeout 600 $TR --master-port 29513 tools/slab_check.py > gpurun_out/slab_check_nccl_$N.log 2>&1; echo "slab_check (nccl) exit $?" >> gpurun_out/slab_check_nccl_$N.log
grep -E "FAIL|PASSED|FAILED|Error|error|exit" gpurun_out/slab_check_nccl_$N.log | head -20
fi
timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_slab_$N.json 2> gpurun_out/bench_slab_$N.err
cat gpurun_out/bench_slab_$N.json; grep -E "Error|error" gpurun_out/bench_slab_$N.err | head -5
MHH_NO_PEER=1 timeout 400 $TR --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_slab_nccl_$N.json 2> gpurun_out/bench_slab_nccl_$N.err
cat gpurun_out/bench_slab_nccl_$N.json
