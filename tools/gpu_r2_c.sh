#!/bin/bash
# Round 2, GPU pass C: fused Poisson path (parity + A/B against the three-kernel version), full parity suite.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fft_roundtrip or pres_2 or full_rk3_step" > gpurun_out/pytest_pres.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_pres.log
tail -12 gpurun_out/pytest_pres.log
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS} > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
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
BENCH_ARGS="--dtype f32" run f32_unfused MHH_PRES_FUSED=0
BENCH_ARGS="--workload 256x256x256" run fused_256
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
du -sh gpurun_out
