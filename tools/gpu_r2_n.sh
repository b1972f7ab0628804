#!/bin/bash
# pass N: ncu source-level captures of the final mom3 (aligned halo) and evisc3 (TMA) kernels
set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-side-configs --workload 512x512x512"
for K in evisc3_kernel mom3_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 9 -c 1 -o gpurun_out/prof_$K -f $B > gpurun_out/ncu_$K.log 2>&1
  ncu -i gpurun_out/prof_$K.ncu-rep --page raw --csv > gpurun_out/final_${K}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$K.ncu-rep --page source --csv > gpurun_out/final_${K}_source.csv 2>/dev/null
  rm -f gpurun_out/prof_$K.ncu-rep
done
ls -la gpurun_out/final_*
