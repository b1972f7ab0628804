#!/bin/bash
# pass R: two-cells-per-thread streaming kernels (rk3, pres_out_rk3): parity, then A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_golden.py -m gpu -q -x -k "rk3 or timeloop or full_rk3_step or drycblles or golden or two_steps or multi_tile" > gpurun_out/pytest_r.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -5 gpurun_out/pytest_r.log | cut -c1-300
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
run r_v1 MHH_STREAM_VEC2=0
run r_v2
BENCH_ARGS="--dtype f32"
run r_v1_f32 MHH_STREAM_VEC2=0
run r_v2_f32
