#!/bin/bash
# Round 2, session 3: (1) the new GPU tests (Diff_tke2 + limiter, Advec_2 inside the fused TMA kernel), (2) the whole GPU suite,
# (3) A/B of the one-sided flux65 form (lib = MHH_UPWIND 1, lib_ab = classic vel*i6 - |vel|*i5) at 512^3 fp64 / fp32,
# (4) drycblles as shipped (swadvec = 2): fused vs two point-wise kernels.  Every step under its own timeout.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tke2.py -q -x > gpurun_out/x_pytest_tke2.log 2>&1; echo "tke2 exit $?"; tail -25 gpurun_out/x_pytest_tke2.log | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_shapes.py -q -x -k "advec2 or drycblles" > gpurun_out/x_pytest_advec2.log 2>&1; echo "advec2 exit $?"; tail -25 gpurun_out/x_pytest_advec2.log | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/x_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/x_pytest_gpu.log | cut -c1-300
B="python bench.py --workload 512x512x512 --no-side-configs --no-cpu-baseline --no-e2e --steps 5 --warmup 3"
for dt in f64 f32; do
  timeout 200 $B --dtype $dt > gpurun_out/x_ab_upwind_$dt.json 2> gpurun_out/x_ab_upwind_$dt.err; echo "upwind $dt exit $?"
  MHH_LIB=$PWD/microhh_b200/lib_ab/libmhhb200.so timeout 200 $B --dtype $dt > gpurun_out/x_ab_classic_$dt.json 2> gpurun_out/x_ab_classic_$dt.err; echo "classic $dt exit $?"
done
timeout 200 $B --swadvec 2 > gpurun_out/x_ab_advec2_fused.json 2> gpurun_out/x_ab_advec2_fused.err; echo "advec2 fused exit $?"
MHH_FUSE_ADVEC2=0 timeout 200 $B --swadvec 2 > gpurun_out/x_ab_advec2_split.json 2> gpurun_out/x_ab_advec2_split.err; echo "advec2 split exit $?"
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/x_ab_*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        k = d.get("kernels_ms_per_step") or {}
        print(p.split("/")[-1], "ms/step", round(d["ms_per_step"], 3), "frac", round(d["whole_step_roofline"]["frac_of_hbm"], 4),
              {n: round(v, 2) for n, v in list(k.items())[:6]}, d.get("post_step_divergence", {}).get("relative_to_umax_over_dx"))
    except Exception as e:
        print(p, "FAILED", e); print(open(p.replace(".json", ".err")).read()[-800:])
PY
