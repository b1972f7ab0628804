#!/bin/bash
# pass U: vectorised rhs loads in the x-forward FFT kernel (parity, A/B), and the Taylor-Green known-answer test
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_taylorgreen.py tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -q -x -k "taylorgreen or pres_2 or full_rk3_step or drycblles or two_steps or wfft or multi_tile" > gpurun_out/pytest_u.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -5 gpurun_out/pytest_u.log | cut -c1-300
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
    print(tag, "%.2f ms/step" % d["ms_per_step"], "finite", d.get("finite"), d.get("post_step_divergence", {}).get("relative_to_umax_over_dx"), " ".join(f"{n.replace('_kernel','')}={v:.2f}" for n, v in list(k.items())[:9]))
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/ab_{tag}.err").read()[-800:])
PY
}
BENCH_ARGS=""
run u_scalar MHH_RHS_VEC=0
run u_vec
BENCH_ARGS="--dtype f32"
run u_scalar_f32 MHH_RHS_VEC=0
run u_vec_f32
