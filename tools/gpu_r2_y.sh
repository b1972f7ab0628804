#!/bin/bash
# Round 2, session 3, second pass: new GPU tests (graph replay, restart IO, tke2 after the tolerance fix), the whole GPU suite,
# smoke, the 512^3 fp32 line with the one-sided flux form, and the side configs with graph replay (small grids).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graph.py tests/test_gpu_io.py tests/test_gpu_tke2.py -q > gpurun_out/y_pytest_new.log 2>&1; echo "new tests exit $?"; tail -30 gpurun_out/y_pytest_new.log | cut -c1-600
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/y_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/y_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/y_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/y_smoke.log | cut -c1-400
timeout 600 python bench.py --workload 512x512x512 --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/y_bench_512_sides.json 2> gpurun_out/y_bench_512_sides.err; echo "bench exit $?"
MHH_GRAPH=0 timeout 300 python bench.py --workload 128x128x128 --no-cpu-baseline --no-side-configs --steps 20 --warmup 3 > gpurun_out/y_bench_128_eager.json 2> gpurun_out/y_bench_128_eager.err; echo "128 exit $?"
python - <<'PY'
import json
for tag in ("512_sides", "128_eager"):
    try:
        d = json.loads(open(f"gpurun_out/y_bench_{tag}.json").read().strip().splitlines()[-1])
        print(tag, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("whole_step_roofline"))
        print("   e2e", d.get("e2e"))
        for o in d.get("other_configs") or []:
            print("   side", o.get("workload")[:60], "ms", o.get("ms_per_step"), "eager", o.get("ms_per_step_eager_profiled"), "replays", o.get("graph_replays"), "frac", o.get("frac_of_hbm"), o.get("error"))
    except Exception as e:
        print(tag, "FAILED", e); print(open(f"gpurun_out/y_bench_{tag}.err").read()[-1500:])
PY
