#!/bin/bash
# 2 GPUs: (1) IPC mapping failure injected on rank 1 -> every rank must fall back to NCCL and the parity check must pass;
# (2) a short bench in the normal (peer) mode.
# The fault injection is compiled out of the shipped library: this script needs a build with the knob,
#   make -C microhh_b200/csrc clean && make -C microhh_b200/csrc -j6 EXTRA=-DMHH_TEST_KNOBS
# (and a plain rebuild afterwards).
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
MHH_FAIL_PEER_RANK=1 timeout 300 $TR --master-port 29511 tools/slab_check.py > gpurun_out/slab_check_fallback.log 2>&1; echo "slab_check (injected IPC failure) exit $?" >> gpurun_out/slab_check_fallback.log
grep -E "FAIL|PASSED|FAILED|Error|error|exit|IPC" gpurun_out/slab_check_fallback.log | head -10
timeout 200 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_slab_2_final.json 2> gpurun_out/bench_slab_2_final.err
cut -c1-200 gpurun_out/bench_slab_2_final.json
