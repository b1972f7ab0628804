#!/bin/bash
# Round 2, multi-GPU pass (run with gpurun --gpus N): slab parity (oracle, bitwise vs single GPU, 4th-order DNS), then bench lines.
set -x
N=${NGPU:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/slab_check.py > gpurun_out/slab_check_n${N}_peer.log 2>&1; echo "slab_check exit $?" >> gpurun_out/slab_check_n${N}_peer.log
grep -E "FAIL|PASSED|FAILED|exit|Error|error" gpurun_out/slab_check_n${N}_peer.log | head -20
MHH_NO_PEER=1 timeout 600 $TR --master-port 29512 tools/slab_check.py > gpurun_out/slab_check_n${N}_nccl.log 2>&1; echo "slab_check exit $?" >> gpurun_out/slab_check_n${N}_nccl.log
grep -E "FAIL|PASSED|FAILED|exit|Error|error" gpurun_out/slab_check_n${N}_nccl.log | head -20
timeout 400 $TR --master-port 29513 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}_strong.json 2> gpurun_out/bench_n${N}_strong.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n${N}_strong.json'))
    print('strong', d['n_gpus'], f"{d['ms_per_step']:.2f} ms/step", d['config']['global_grid'], d['nvlink'], 'e2e', d['e2e'] and d['e2e']['value'], 'div', d.get('post_step_divergence'))
    print(' '.join(f"{k.replace('_kernel','')}={v:.2f}" for k,v in d['kernels_ms_per_step'].items()))
except Exception as e:
    print('strong FAILED', e, open('gpurun_out/bench_n${N}_strong.err').read()[-1500:])
PY
timeout 400 $TR --master-port 29514 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --scaling weak --workload 512x512x512 > gpurun_out/bench_n${N}_weak.json 2> gpurun_out/bench_n${N}_weak.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n${N}_weak.json'))
    print('weak', d['n_gpus'], f"{d['ms_per_step']:.2f} ms/step", d['config']['global_grid'], d['nvlink'], 'div', d.get('post_step_divergence'))
    print(' '.join(f"{k.replace('_kernel','')}={v:.2f}" for k,v in d['kernels_ms_per_step'].items()))
except Exception as e:
    print('weak FAILED', e, open('gpurun_out/bench_n${N}_weak.err').read()[-1500:])
PY
