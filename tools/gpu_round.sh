#!/bin/bash
# One single-GPU box pass: parity tests, bench lines, ncu launch list + full captures exported as CSV
# (the .ncu-rep files are too big for the gpurun_out/ return channel and are deleted after export).
set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_512_f64.json 2> gpurun_out/bench_512_f64.err
cat gpurun_out/bench_512_f64.json
timeout 600 python bench.py --steps 5 --warmup 3 --dtype f32 --no-cpu-baseline > gpurun_out/bench_512_f32.json 2> gpurun_out/bench_512_f32.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 python tools/bench_o4.py > gpurun_out/bench_o4.json 2> gpurun_out/bench_o4.err
cat gpurun_out/bench_o4.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_512_f64.csv $B > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mom3_kernel|evisc_tile" -s 6 -c 2 -o gpurun_out/tend_full -f $B > gpurun_out/ncu_tend.log 2>&1
ncu -i gpurun_out/tend_full.ncu-rep --page raw --csv > gpurun_out/tend_full_raw.csv 2>/dev/null
rm -f gpurun_out/tend_full.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:"wfft|tdma_solve|pres_out_rk3|rk3_kernel|cyclic" -s 30 -c 8 -o gpurun_out/pres_full -f $B > gpurun_out/ncu_pres.log 2>&1
ncu -i gpurun_out/pres_full.ncu-rep --page raw --csv > gpurun_out/pres_full_raw.csv 2>/dev/null
rm -f gpurun_out/pres_full.ncu-rep
du -sh gpurun_out
