#!/bin/bash
# pass O: evisc3 with the th planes staged by TMA and z0 in shared memory -- parity, then timing
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -q -x -k "diff_smag2 or evisc or fused_tendencies or drycblles or full_rk3_step" > gpurun_out/pytest_o.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -5 gpurun_out/pytest_o.log | cut -c1-300
[ $rc -eq 124 ] && exit 1
MHH_EVISC3_NPL=1 timeout 600 python -m pytest tests/test_gpu_shapes.py -m gpu -q -x -k "fused_tendencies" > gpurun_out/pytest_o1.log 2>&1
rc=$?; echo "pytest npl1 exit $rc"; tail -3 gpurun_out/pytest_o1.log | cut -c1-300
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
    print(tag, "%.2f ms/step" % d["ms_per_step"], "finite", d.get("finite"), " ".join(f"{n.replace('_kernel','')}={v:.2f}" for n, v in list(k.items())[:4]))
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/ab_{tag}.err").read()[-800:])
PY
}
BENCH_ARGS=""
run o_n2m2
run o_n1m3 MHH_EVISC3_NPL=1 MHH_EVISC3_MB=3
run o_n1m4 MHH_EVISC3_NPL=1 MHH_EVISC3_MB=4
run o_n1m2 MHH_EVISC3_NPL=1 MHH_EVISC3_MB=2
BENCH_ARGS="--dtype f32"
run o_n2m4_f32
run o_n1m4_f32 MHH_EVISC3_NPL=1 MHH_EVISC3_MB=4
run o_n2m3_f32 MHH_EVISC3_MB=3
