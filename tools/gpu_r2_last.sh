#!/bin/bash
# last pass of the round: the default bench line (side lines with the TMA scalar kernel) and the fp32 line
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/l_bench_default.json 2> gpurun_out/l_bench_default.err; echo "bench exit $?"
timeout 300 python bench.py --workload 512x512x512 --dtype f32 --no-side-configs --no-cpu-baseline > gpurun_out/l_bench_512_f32.json 2> gpurun_out/l_bench_512_f32.err; echo "f32 exit $?"
python - <<'PY'
import json
for tag in ("default", "512_f32"):
    try:
        d = json.loads(open(f"gpurun_out/l_bench_{tag}.json").read().strip().splitlines()[-1])
        print(tag, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
        print("   roofline", d.get("roofline"), "whole", d.get("whole_step_roofline"))
        print("   e2e", d.get("e2e")); print("   cpu", d.get("cpu_baseline"))
        for o in d.get("other_configs") or []:
            print("   side", (o.get("workload") or "")[:70], "ms", o.get("ms_per_step"), "eager", o.get("ms_per_step_eager_profiled"), "second", o.get("ms_per_step_plain_second_pass"), "frac", o.get("frac_of_hbm"), o.get("error"))
            print("        ", {k: round(v, 3) for k, v in (o.get("kernels_ms_per_step") or {}).items()})
    except Exception as e:
        print(tag, "FAILED", e); print(open(f"gpurun_out/l_bench_{tag}.err").read()[-1200:])
PY
