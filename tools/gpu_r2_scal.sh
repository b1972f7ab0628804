#!/bin/bash
# A/B of the TMA-staged scalar kernel (scalar group of mom3 alone) against the cp.async tile kernel, then the whole GPU suite.
mkdir -p gpurun_out
for cfg in "512x512x256 f32" "256x256x256 f64" "512x512x256 f64"; do
  set -- $cfg
  for sw in 1 0; do
    MHH_SCAL_TMA=$sw timeout 200 python tools/tke2_bench.py --grid $1 --dtype $2 > gpurun_out/s_scal_${1}_${2}_tma$sw.json 2> gpurun_out/s_scal_${1}_${2}_tma$sw.err
  done
done
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/s_scal_*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1]); k = d["kernels_ms_per_step"]
        print(p.split("/")[-1], "ms/step", round(d["ms_per_step"], 3), {n: round(v, 3) for n, v in k.items() if n.startswith("scal") or n.startswith("mom3")}, d["finite"])
    except Exception as e:
        print(p, "FAILED", e); print(open(p.replace(".json", ".err")).read()[-600:])
PY
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/s_pytest_gpu.log | cut -c1-400
