#!/bin/bash
# Round 2, GPU pass F: mom3 with the aligned halo (igc = 4), occupancy fixes of the fused Poisson kernels, unit-stride rhs loads.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shapes.py tests/test_gpu_parity.py -m gpu -q -x -k "multi_tile or fused_tend or drycblles or fft_roundtrip or pres_2 or full_rk3 or two_steps" > gpurun_out/pytest_f.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_f.log
tail -8 gpurun_out/pytest_f.log | cut -c1-300
if [ $rc -eq 124 ]; then echo "HANG: stopping"; exit 1; fi
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
BENCH_ARGS="--igc 3" run f64_igc3
BENCH_ARGS="--igc 4" run f64_igc4
BENCH_ARGS="--igc 4" run f64_igc4_ty4 MHH_TILE3_Y=4
BENCH_ARGS="--dtype f32" run f32_igc4
du -sh gpurun_out
