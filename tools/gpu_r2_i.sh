#!/bin/bash
# Round 2, GPU pass I: two-sided factorisation in the fused y kernels (parity + bench).
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fft_roundtrip" > gpurun_out/pytest_i0.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_i0.log
tail -5 gpurun_out/pytest_i0.log | cut -c1-300
if [ $rc -eq 124 ]; then echo "HANG: stopping"; exit 1; fi
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_fft_vs_cufft.py -m gpu -q -x -k "pres_2 or full_rk3 or two_steps or wfft or generic_fft or drycblles or cufft" > gpurun_out/pytest_i.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_i.log
tail -8 gpurun_out/pytest_i.log | cut -c1-300
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
run twist
BENCH_ARGS="--dtype f32" run twist_f32
