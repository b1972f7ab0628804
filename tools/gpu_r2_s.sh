#!/bin/bash
# pass S: fp32 mom3 compiled for two resident CTAs per SM: parity (fp32 cases), then timing
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_golden.py -m gpu -q -x -k "float32 or f32 or fp32" > gpurun_out/pytest_s.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -5 gpurun_out/pytest_s.log | cut -c1-300
[ $rc -eq 124 ] && exit 1
for tag in s_f32 s_f64; do
  A=""; [ $tag = s_f32 ] && A="--dtype f32"
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-side-configs --workload 512x512x512 $A > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
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
done
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workload 512x512x512 --dtype f32 > gpurun_out/ab_s_side.json 2> gpurun_out/ab_s_side.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/ab_s_side.json").read().strip().splitlines()[-1])
for o in d.get("other_configs") or []:
    print("side", o.get("workload"), o.get("ms_per_step"), o.get("frac_of_hbm"), o.get("error"), o.get("kernels_ms_per_step"))
PY
