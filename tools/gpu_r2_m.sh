#!/bin/bash
# pass M: what bounds the TMA eddy-viscosity kernel -- planes in flight (ring depth) vs. tile width
set -x
mkdir -p gpurun_out
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
    print(tag, "%.2f ms/step" % d["ms_per_step"], "finite", d.get("finite"), " ".join(f"{n.replace('_kernel','')}={v:.2f}" for n, v in list(k.items())[:4]))
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/ab_{tag}.err").read()[-800:])
PY
}
BENCH_ARGS=""
run m_n1r6m4 MHH_EVISC3_NPL=1 MHH_EVISC3_RING=6 MHH_EVISC3_MB=4
run m_n1r8m3 MHH_EVISC3_NPL=1 MHH_EVISC3_RING=8 MHH_EVISC3_MB=3
run m_n2r6m2 MHH_EVISC3_NPL=2 MHH_EVISC3_RING=6 MHH_EVISC3_MB=2
run m_n4r4 MHH_EVISC3_NPL=4 MHH_EVISC3_RING=4
run m_n4r6 MHH_EVISC3_NPL=4 MHH_EVISC3_RING=6
