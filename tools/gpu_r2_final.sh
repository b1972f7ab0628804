#!/bin/bash
# Round 2, final single-GPU validation: smoke, the whole GPU suite, the default bench line and the reference arm, the ncu
# launch list of the bench command.  Every step under its own timeout.
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/final_smoke.log | cut -c1-400
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/final_pytest_gpu.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -6 gpurun_out/final_pytest_gpu.log | cut -c1-300
[ $rc -eq 124 ] && exit 1
timeout 900 python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference exit $?"
timeout 300 python bench.py --workload 512x512x512 --dtype f32 --no-side-configs --no-cpu-baseline > gpurun_out/final_bench_512_f32.json 2> gpurun_out/final_bench_512_f32.err; echo "f32 exit $?"
python - <<'PY'
import json
for tag in ("default", "reference", "512_f32"):
    try:
        d = json.loads(open(f"gpurun_out/final_bench_{tag}.json").read().strip().splitlines()[-1])
        print(tag, {k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "steps", "gpu_launches", "clocks")}, (d.get("config") or {}).get("workload"))
        print("   roofline", d.get("roofline"), "whole", d.get("whole_step_roofline"))
        print("   e2e", d.get("e2e")); print("   cpu", d.get("cpu_baseline"))
        for o in d.get("other_configs") or []:
            print("   side", o.get("workload"), o.get("ms_per_step"), o.get("frac_of_hbm"), o.get("error"))
    except Exception as e:
        print(tag, "FAILED", e); print(open(f"gpurun_out/final_bench_{tag}.err").read()[-1200:])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/final_ncu_bench.log 2>&1; echo "ncu exit $?"
wc -l gpurun_out/final_launches.csv
