#!/bin/bash
# Round 2, GPU pass B: mom4 parity subset, A/B of the momentum kernel variants (and of nvcc -split-compile), full parity suite, ncu.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests/test_gpu_shapes.py -m gpu -q -x -k "multi_tile or fused_tend" > gpurun_out/pytest_mom4.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mom4.log
tail -5 gpurun_out/pytest_mom4.log
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS} > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/ab_{tag}.json'))
    print(tag, f"{d['ms_per_step']:.2f} ms/step", d['clocks'], ' '.join(f"{k.replace('_kernel','')}={v:.2f}" for k,v in d['kernels_ms_per_step'].items()))
except Exception as e:
    print(tag, 'FAILED', e, open(f'gpurun_out/ab_{tag}.err').read()[-500:])
PY
}
run mom3 MHH_MOM=3
run mom4_w2 MHH_MOM=4 MHH_TILE4_W=2
run mom4_w3 MHH_MOM=4 MHH_TILE4_W=3
run mom4_w2_pf2 MHH_MOM=4 MHH_TILE4_W=2 MHH_PREFETCH=2
BENCH_ARGS="--dtype f32" run f32_mom4_w3 MHH_MOM=4 MHH_TILE4_W=3
BENCH_ARGS="--dtype f32" run f32_mom4_w2 MHH_MOM=4 MHH_TILE4_W=2
BENCH_ARGS="--dtype f32 --igc 3" run f32_igc3
if [ -f microhh_b200/lib/libmhhb200_nosplit.so ]; then
  cp microhh_b200/lib/libmhhb200.so /tmp/split.so; cp microhh_b200/lib/libmhhb200_nosplit.so microhh_b200/lib/libmhhb200.so
  run nosplit_mom3 MHH_MOM=3
  run nosplit_mom4_w2 MHH_MOM=4 MHH_TILE4_W=2
  cp /tmp/split.so microhh_b200/lib/libmhhb200.so
fi
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mom4_kernel" -s 9 -c 1 -o gpurun_out/mom4_full -f $B > gpurun_out/ncu_mom4.log 2>&1
ncu -i gpurun_out/mom4_full.ncu-rep --page raw --csv > gpurun_out/mom4_full_raw.csv 2>/dev/null
ncu -i gpurun_out/mom4_full.ncu-rep --page source --csv > gpurun_out/mom4_full_source.csv 2>/dev/null
rm -f gpurun_out/mom4_full.ncu-rep
du -sh gpurun_out
