#!/bin/bash
# A/B on one box: mom3 with the face velocities as sums (lib) vs as interp2 (lib_ab = the previous tile3_kernels.cuh), 512^3 fp64,
# interleaved twice so that clock drift shows.
mkdir -p gpurun_out
B="python bench.py --workload 512x512x512 --no-side-configs --no-cpu-baseline --no-e2e --steps 5 --warmup 3"
for r in 1 2; do
  timeout 200 $B > gpurun_out/ab2_sum2_$r.json 2> gpurun_out/ab2_sum2_$r.err
  MHH_LIB=$PWD/microhh_b200/lib_ab/libmhhb200.so timeout 200 $B > gpurun_out/ab2_interp2_$r.json 2> gpurun_out/ab2_interp2_$r.err
done
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/ab2_*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1]); k = d["kernels_ms_per_step"]
        print(p.split("/")[-1], "ms/step", round(d["ms_per_step"], 3), "mom3", round(k["mom3_kernel"], 3), "evisc3", round(k["evisc3_kernel"], 3), "x_fwd", round(k["fft_x_forward_kernel"], 3), d["clocks"])
    except Exception as e:
        print(p, "FAILED", e)
PY
