#!/bin/bash
# pass K: TMA-staged eddy-viscosity kernel -- parity, then A/B timing against the cp.async kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_golden.py -m gpu -q -x -k "diff_smag2 or evisc or fused_tendencies or drycblles or full_rk3_step or golden or pres_2" > gpurun_out/pytest_k.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -8 gpurun_out/pytest_k.log | cut -c1-300
[ $rc -eq 124 ] && exit 1
run() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-side-configs --workload 512x512x512 $BENCH_ARGS > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  [ $? -eq 124 ] && { echo "TIMEOUT $tag"; exit 1; }
  python - $tag <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{tag}.json").read().strip().splitlines()[-1])
    k = d["kernels_ms_per_step"]
    print(tag, "%.2f ms/step" % d["ms_per_step"], " ".join(f"{n.replace('_kernel','')}={v:.2f}" for n, v in list(k.items())[:6]))
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/ab_{tag}.err").read()[-800:])
PY
}
BENCH_ARGS=""
run old MHH_EVISC_TMA=0
run tma3 MHH_EVISC_TMA=1
run tma2 MHH_EVISC_TMA=1 MHH_EVISC3_MB=2
run tma4 MHH_EVISC_TMA=1 MHH_EVISC3_MB=4
BENCH_ARGS="--dtype f32"
run old_f32 MHH_EVISC_TMA=0
run tma3_f32 MHH_EVISC_TMA=1
run tma4_f32 MHH_EVISC_TMA=1 MHH_EVISC3_MB=4
