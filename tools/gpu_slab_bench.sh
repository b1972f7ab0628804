#!/bin/bash
# bench only at N ranks (gpurun --gpus N)
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_slab_$N.json 2> gpurun_out/bench_slab_$N.err
cat gpurun_out/bench_slab_$N.json; grep -E "Error|error" gpurun_out/bench_slab_$N.err | head -5
