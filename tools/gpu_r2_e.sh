#!/bin/bash
# Round 2, GPU pass E: re-run the tests fixed after pass D; ncu source-level captures of the four kernels furthest from the roofline.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_shapes.py tests/test_gpu_surface.py tests/test_fft_vs_cufft.py -m gpu -q > gpurun_out/pytest_fixed.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_fixed.log
tail -25 gpurun_out/pytest_fixed.log | cut -c1-250
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-side-configs --workload 512x512x512"
for K in mom3_kernel evisc_tile_kernel p2_y_forward_kernel p2_y_backward_kernel p2_x_backward_kernel p2_x_forward_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 9 -c 1 -o gpurun_out/prof_$K -f $B > gpurun_out/ncu_$K.log 2>&1
  ncu -i gpurun_out/prof_$K.ncu-rep --page raw --csv > gpurun_out/${K}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$K.ncu-rep --page source --csv > gpurun_out/${K}_source.csv 2>/dev/null
  rm -f gpurun_out/prof_$K.ncu-rep
done
du -sh gpurun_out
