#!/bin/bash
# A/B runs of bench.py under different MHH_* knobs; prints ms/step and the per-kernel split.
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS} > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/ab_{tag}.json'))
    print(tag, f"{d['ms_per_step']:.2f} ms/step", ' '.join(f"{k.replace('_kernel','')}={v:.2f}" for k,v in d['kernels_ms_per_step'].items()))
except Exception as e:
    print(tag, 'FAILED', e, open(f'gpurun_out/ab_{tag}.err').read()[-500:])
PY
}
