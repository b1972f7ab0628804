#!/bin/bash
# Round 2, GPU pass G: forcing tests; A/B of the two rhs loaders of the x-forward kernel.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_forcing.py -m gpu -q > gpurun_out/pytest_g.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_g.log
tail -12 gpurun_out/pytest_g.log | cut -c1-300
run() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-side-configs --workload 512x512x512 ${BENCH_ARGS} > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/ab_{tag}.json'))
    print(tag, f"{d['ms_per_step']:.2f} ms/step", d['clocks'], ' '.join(f"{k.replace('_kernel','')}={v:.2f}" for k,v in d['kernels_ms_per_step'].items()))
except Exception as e:
    print(tag, 'FAILED', e, open(f'gpurun_out/ab_{tag}.err').read()[-700:])
PY
}
run pair
BENCH_ARGS="--dtype f32" run pair_f32
cp microhh_b200/lib/libmhhb200.so /tmp/main.so; cp microhh_b200/lib/libmhhb200_unroll4.so microhh_b200/lib/libmhhb200.so
run unit4
BENCH_ARGS="--dtype f32" run unit4_f32
cp /tmp/main.so microhh_b200/lib/libmhhb200.so
