#!/bin/bash
# pass W: scalar RK3 update forked onto a side stream under the pressure solve: parity, then A/B
set -x
mkdir -p gpurun_out
true
rc=0
[ $rc -eq 124 ] && exit 1
run() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-side-configs --workload 512x512x512 $BENCH_ARGS > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  [ $? -eq 124 ] && { echo "TIMEOUT $tag"; exit 1; }
  python - $tag <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{tag}.json").read().strip().splitlines()[-1])
    k = d["kernels_ms_per_step"]
    print(tag, "%.2f ms/step" % d["ms_per_step"], "finite", d.get("finite"), d.get("post_step_divergence", {}).get("relative_to_umax_over_dx"), " ".join(f"{n.replace('_kernel','')}={v:.2f}" for n, v in list(k.items())[:10]))
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/ab_{tag}.err").read()[-800:])
PY
}
BENCH_ARGS=""
run w_seq MHH_OVERLAP=0
run w_ovl
BENCH_ARGS="--dtype f32"
run w_seq_f32 MHH_OVERLAP=0
run w_ovl_f32
