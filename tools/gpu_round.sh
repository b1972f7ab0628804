#!/bin/bash
# One single-GPU box pass: parity tests, bench line, ncu launch list + full captures exported as CSV
# (the .ncu-rep files are too big for the gpurun_out/ return channel and are deleted after export).
set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt
free -g >> gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_512_f64.json 2> gpurun_out/bench_512_f64.err
cat gpurun_out/bench_512_f64.json
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_512_f64.csv $B > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mom3_kernel|evisc_tile" -s 6 -c 2 -o gpurun_out/tend_full -f $B > gpurun_out/ncu_tend.log 2>&1
ncu -i gpurun_out/tend_full.ncu-rep --page raw --csv > gpurun_out/tend_full_raw.csv 2>/dev/null
ncu -i gpurun_out/tend_full.ncu-rep --page details --csv > gpurun_out/tend_full_details.csv 2>/dev/null
ncu -i gpurun_out/tend_full.ncu-rep --page source --csv --kernel-name regex:evisc > gpurun_out/evisc_source.csv 2>/dev/null
ncu -i gpurun_out/tend_full.ncu-rep --page source --csv --kernel-name regex:mom3 > gpurun_out/mom3_source.csv 2>/dev/null
rm -f gpurun_out/tend_full.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:"wfft|tdma_solve|pres_out_rk3|rk3_kernel|cyclic" -s 30 -c 8 -o gpurun_out/pres_full -f $B > gpurun_out/ncu_pres.log 2>&1
ncu -i gpurun_out/pres_full.ncu-rep --page raw --csv > gpurun_out/pres_full_raw.csv 2>/dev/null
ncu -i gpurun_out/pres_full.ncu-rep --page details --csv > gpurun_out/pres_full_details.csv 2>/dev/null
rm -f gpurun_out/pres_full.ncu-rep
gzip -f gpurun_out/*_source.csv
du -sh gpurun_out; ls -la gpurun_out
