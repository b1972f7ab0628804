#!/bin/bash
# pass J: Advec_4m parity (kernels, o4 steps incl. moser180 shape as shipped) + the moser180-shaped bench side line
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -q -x -k "advec_4 or order4 or moser180" > gpurun_out/pytest_j.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -8 gpurun_out/pytest_j.log | cut -c1-300
[ $rc -eq 124 ] && exit 1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workload 512x512x512 > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_j.json").read().strip().splitlines()[-1])
    print("main", d["ms_per_step"], d["roofline"]["frac"])
    for o in d.get("other_configs", []):
        print(o.get("workload"), o.get("ms_per_step"), o.get("frac_of_hbm"), o.get("error"), o.get("kernels_ms_per_step"))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_j.err").read()[-1500:])
PY
