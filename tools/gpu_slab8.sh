#!/bin/bash
set -x
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/slab_check.py > gpurun_out/slab_check_$N.log 2>&1; echo "slab_check (peer) exit $?" >> gpurun_out/slab_check_$N.log
grep -E "FAIL|PASSED|FAILED|Error|error|exit" gpurun_out/slab_check_$N.log | head -20
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_slab_$N.json 2> gpurun_out/bench_slab_$N.err
cat gpurun_out/bench_slab_$N.json; grep -E "Error|error" gpurun_out/bench_slab_$N.err | head -5
