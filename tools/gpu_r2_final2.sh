#!/bin/bash
# Round 2, final single-GPU validation of the last session: smoke, the whole GPU suite, the default bench line and the reference
# arm, the 512^3 fp32 line, the tke2 side measurement, the ncu launch list of the bench command and full captures of the new
# kernel variants.  Every step under its own timeout.
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/z_smoke.log | cut -c1-400
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/z_pytest_gpu.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -6 gpurun_out/z_pytest_gpu.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/z_bench_default.json 2> gpurun_out/z_bench_default.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/z_bench_reference.json 2> gpurun_out/z_bench_reference.err; echo "reference exit $?"
timeout 300 python bench.py --workload 512x512x512 --dtype f32 --no-side-configs --no-cpu-baseline > gpurun_out/z_bench_512_f32.json 2> gpurun_out/z_bench_512_f32.err; echo "f32 exit $?"
timeout 300 python tools/tke2_bench.py > gpurun_out/z_tke2_256_f64.json 2> gpurun_out/z_tke2.err; echo "tke2 exit $?"; cat gpurun_out/z_tke2_256_f64.json | cut -c1-900
python - <<'PY'
import json
for tag in ("default", "reference", "512_f32"):
    try:
        d = json.loads(open(f"gpurun_out/z_bench_{tag}.json").read().strip().splitlines()[-1])
        print(tag, {k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "steps", "gpu_launches", "clocks")}, (d.get("config") or {}).get("workload"))
        print("   roofline", d.get("roofline"), "whole", d.get("whole_step_roofline"))
        print("   e2e", d.get("e2e")); print("   cpu", d.get("cpu_baseline"))
        print("   kernels", d.get("kernels_ms_per_step"))
        for o in d.get("other_configs") or []:
            print("   side", (o.get("workload") or "")[:70], "ms", o.get("ms_per_step"), "eager", o.get("ms_per_step_eager_profiled"), "frac", o.get("frac_of_hbm"), o.get("error"))
    except Exception as e:
        print(tag, "FAILED", e); print(open(f"gpurun_out/z_bench_{tag}.err").read()[-1200:])
PY
KS='regex:^(mom3|evisc|scal_tile|p2_|wfft|fft_|tdma|hdma|pres|rk3|cyclic|ghost|halo|reduce|o2_|o4|surface|buffer|body|coriolis|mean_uut|scalar_forcing|tke2|limiter)'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KS" -c 400 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/z_ncu_bench.log 2>&1; echo "ncu exit $?"
wc -l gpurun_out/z_launches.csv
# full captures (one launch each, after the warm-up launches): Advec_2 variant of mom3, the final 2i5 mom3, tke2_visc
B="python bench.py --workload 512x512x512 --no-side-configs --no-cpu-baseline --no-e2e --steps 1 --warmup 1"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mom3_kernel -s 4 -c 1 -o gpurun_out/z_mom3_advec2 -f $B --swadvec 2 > gpurun_out/z_ncu_a.log 2>&1; echo "ncu advec2 exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mom3_kernel -s 4 -c 1 -o gpurun_out/z_mom3_2i5 -f $B > gpurun_out/z_ncu_b.log 2>&1; echo "ncu 2i5 exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tke2_visc_kernel -s 4 -c 1 -o gpurun_out/z_tke2_visc -f python tools/tke2_bench.py --steps 1 > gpurun_out/z_ncu_c.log 2>&1; echo "ncu tke2 exit $?"
# the reports themselves (~40 MB each) would push gpurun_out/ over the 64 MiB that travel back: keep the CSV pages only
for n in z_mom3_advec2 z_mom3_2i5 z_tke2_visc; do
  if [ -f gpurun_out/$n.ncu-rep ]; then
    ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/${n}_raw.csv 2>/dev/null
    ncu -i gpurun_out/$n.ncu-rep --page source --csv > gpurun_out/${n}_source.csv 2>/dev/null
    rm -f gpurun_out/$n.ncu-rep
  fi
done
ls -la gpurun_out/z_*_raw.csv gpurun_out/z_*_source.csv; du -sh gpurun_out
