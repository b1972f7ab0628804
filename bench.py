#!/usr/bin/env python
"""
bench.py -- grid-point RK3 steps/s of the dynamical-core hot path (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W            our arm (CUDA, libmhhb200.so)
  python bench.py --impl reference ...                      the reference's own CPU kernels on host cores

One "step" = one full RK3 time step (3 sub-steps: cyclic + vertical ghost cells, eddy viscosity,
advection + diffusion + buoyancy tendencies, FFT/tridiagonal pressure solve, pressure correction
+ RK3 update) of a drycblles-shaped LES (advec_2i5 + diff_smag2 + pres_2 + thermo_dry, S = 1
scalar) on a synthetic grid.  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

# NCCL prints its version banner to stdout when NCCL_DEBUG is VERSION/WARN; keep stdout for the ONE JSON line
# (NCCL honours NCCL_DEBUG_FILE only above the VERSION level)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

METRIC = "grid-point RK3 steps/s"
UNIT = "grid-point-steps/s"


def parse_workload(s):
    it, jt, kt = (int(x) for x in s.lower().split("x"))
    return it, jt, kt


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm)//2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


def algorithmic_bytes_per_point_substep(S, B):
    """(41 + 7 S) * sizeof(TF) bytes per grid point per sub-step."""
    return (41 + 7*S)*B


# ----------------------------------------------------------------------------------------------
ALG_PASSES = {  # algorithmic array passes per launch, in units of N*B bytes
    "tend_uvw_kernel": 4 + 6, "tend_s_kernel": 5 + 2, "evisc_kernel": 5,
    # z-marching tile kernels: R u,v,w,evisc,th + RMW ut,vt,wt | R s,u,v,w,evisc + RMW st | R u,v,w,th + W evisc
    # mom3 carries scalar 0 as a fourth warp group: R u,v,w,evisc,th + RMW ut,vt,wt,tht
    "mom_tile_kernel": 5 + 6, "mom3_kernel": 5 + 8, "mom3_kernel_advec2": 5 + 8, "scal_tile_kernel": 5 + 2, "evisc_tile_kernel": 5, "evisc3_kernel": 5,
    "fft_x_forward_kernel": 7, "fft_y_forward_kernel": 2, "fft_y_backward_kernel": 2,
    "tdma_solve_kernel": 2, "fft_x_backward_kernel": 2, "pres_out_rk3_kernel": 13, "rk3_kernel": 4,
    # Pres_2 version 2: y transform fused with the Thomas sweeps (R + W of the spectral array each; the pivot table is overhead)
    "fft_y_tdma_forward_kernel": 2, "tdma_fft_y_backward_kernel": 2,
}


def decompose(workload, world, scaling):
    """Global grid and per-rank slab.  strong (default): the named grid is THE grid at every N (BASELINE configs[4]:
    the 1024^3 dry CBL over 1/2/4/8 B200).  weak: every GPU keeps the named number of points (y, then x, then z doubles)."""
    itot, jtot, ktot = parse_workload(workload)
    if scaling == "weak":
        fx, fy, fz = {1: (1, 1, 1), 2: (1, 2, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(world, (1, world, 1))
        itot, jtot, ktot = itot*fx, jtot*fy, ktot*fz
    return itot, jtot, ktot


def timed_steps(torch, dist, world, dyc, f, dt, steps, warmup, ctx, profile=True):
    """profile=True: per-kernel CUDA events inside the timed region (every step runs eagerly); False: the plain call a user
    makes, which on one GPU replays the CUDA graph of the step from the third call on."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warmup):
        dyc.step(f, dt)
    barrier()
    l0 = ctx.launch_count
    if not profile:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            dyc.step(f, dt)
        e1.record()
        barrier()
        return e0.elapsed_time(e1), {}, ctx.launch_count - l0
    ctx.profile_start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        dyc.step(f, dt)
    e1.record()
    barrier()
    prof = ctx.profile_stop()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, prof, ctx.launch_count - l0


def side_config(torch, D, GridData, fill_fields_device, name, shape, dtype, S, peaks, steps, order=2, dt=1.0, device=0, swadvec="2i5"):
    """One more single-GPU configuration of BASELINE.json's list, reported beside the main line (N = 1 only)."""
    it, jt, kt = shape
    B = np.dtype(dtype).itemsize
    try:
        if order == 4:
            gd = GridData(it, jt, kt, 2*np.pi, np.pi, 2., 3, 3, 3, dtype, order=4)
        else:
            gd = GridData(it, jt, kt, 25.*it, 25.*jt, 25.*kt, 4, 3, 1, dtype)
        ctx = D.Context(gd, device)
        scal = ["th"] + [f"s{n}" for n in range(1, S)] if S > 0 else ["th"]
        f = D.Fields(ctx, None, scalars=scal, visc=1e-5 if order == 2 else 1e-3, svisc=1e-5 if order == 2 else 1e-3)
        prof1d = fill_fields_device(f, gd, noise=0.01)
        ctx.set_basestate(prof1d["rhoref"], prof1d["rhorefh"], prof1d["thref"], prof1d["threfh"])
        if order == 4:
            prm = D.make_params(swadvec="4m", swdiff="4", swthermo=None, surface_model=False, mbcbot=0, mbctop=0, ns=len(scal))
            dt = 1e-3
        else:
            prm = D.make_params(swadvec=swadvec, ns=len(scal))
        dyc = D.Dycore(ctx, prm)
        ms_eager, prof, launches = timed_steps(torch, None, 1, dyc, f, dt, steps, 3, ctx)             # eager, per-kernel events
        r0 = ctx.graph_replays
        ms, _, _ = timed_steps(torch, None, 1, dyc, f, dt, steps, 3, ctx, profile=False)              # the user's call: graph replay
        replays = ctx.graph_replays - r0
        ms_second = ms
        if replays == 0:
            ms = ms_eager       # no graph on this grid: both passes run the same eager step; the first one is the line, as before
        npts = it*jt*kt
        ns = len(scal)
        passes = (35 + 7*ns) if order == 4 else (41 + 7*ns)        # no eddy-viscosity stage in the DNS configuration
        bytes_step = 3*passes*B*npts
        out = {"workload": name, "grid": f"{it}x{jt}x{kt}", "dtype": "f64" if dtype == np.float64 else "f32", "scalars": ns,
               "ms_per_step": ms/steps, "ms_per_step_eager_profiled": ms_eager/steps, "ms_per_step_plain_second_pass": ms_second/steps,
               "graph_replays": replays,
               "value": npts*steps/(ms*1e-3), "unit": UNIT,
               "algorithmic_bytes_per_point_step": 3*passes*B,
               "frac_of_hbm": bytes_step/(ms/steps*1e-3)/1e9/peaks["hbm_gbs"],
               "finite": bool(torch.isfinite(f["u"]).all().item()), "gpu_launches": launches,
               "kernels_ms_per_step": {k: v["ms"]/steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]}}
        ctx.close()
        del f
        torch.cuda.empty_cache()
        return out
    except Exception as ex:     # a side line must never cost the main one
        return {"workload": name, "error": str(ex)[:300]}


def thermo_side(thermo, grid, dtype, steps):
    """Side lines of the thermo couplings (tools/thermo_bench.py: the step with Thermo_moist / Thermo_buoy registered into the fused
    sub-step, per-kernel times of the thermo kernels); like every side line it must never cost the main one."""
    try:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
        import thermo_bench
        return thermo_bench.run(thermo, grid, dtype, steps)
    except Exception as ex:
        return {"workload": f"thermo side line {thermo} {grid} {dtype}", "error": str(ex)[:300]}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from microhh_b200 import dycore as D
    from microhh_b200.grid import GridData
    from microhh_b200.synthetic import make_case, fill_fields_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dtype = np.float64 if args.dtype == "f64" else np.float32
    B = np.dtype(dtype).itemsize
    S = 1
    itot_g, jtot_g, ktot = decompose(args.workload, world, args.scaling)
    # the adapters ask Grid for igc = 4 (Grid::set_minimum_ghost_cells): aligned pairs in the TMA-staged kernels, and for
    # USESP a row pitch that is a multiple of 16 B; --igc 3 runs the reference's minimum for advec_2i5
    igc = args.igc if args.igc else 4
    gd = GridData(itot_g, jtot_g, ktot, 25.*itot_g, 25.*jtot_g, 25.*ktot, igc, 3, 1, dtype, npy=world, mpicoordy=rank)
    ctx = D.Context(gd, local_rank)
    f = D.Fields(ctx, None)
    prof1d = fill_fields_device(f, gd, seed=2, noise=0.01)          # generated on the device: 1024^3 never exists on the host
    ctx.set_basestate(prof1d["rhoref"], prof1d["rhorefh"], prof1d["thref"], prof1d["threfh"])
    prm = D.make_params(swadvec=args.swadvec)
    dyc = D.Dycore(ctx, prm)
    dt = args.dt
    npts = gd.npoints // world        # per GPU

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, prof, launches = timed_steps(torch, dist, world, dyc, f, dt, args.steps, args.warmup, ctx)
    if rank == 0:
        sampler.stop_flag.set(); sampler.join(timeout=2)
    finite = bool(torch.isfinite(f["u"]).all().item())
    # size-independent property at the full size: after the pressure correction the velocity is divergence-free to
    # rounding (Pres_2::check_divergence on the cyclic-filled fields; relative to |u|max/dx).  On slabs this is collective
    # (halo rows over NVLink, max all-reduced: the same calls tools/slab_check.py validates) and checks the distributed
    # transposes of the solve at the benchmarked size.
    post_div = None
    try:
        bc = D.Boundary_cyclic(ctx)
        for n in ("u", "v", "w"):
            bc.exec(f[n])
        div = D.Pres(ctx).check_divergence(f)
        umax_t = f["u"].abs().max().to(torch.float64).reshape(1)
        if world > 1:
            dist.all_reduce(umax_t, op=dist.ReduceOp.MAX)
        umax = float(umax_t.item())
        post_div = {"max_abs_divergence": div, "relative_to_umax_over_dx": div/(umax/float(gd.dx))}
    except Exception as ex:      # a diagnostic must never cost the bench line
        post_div = {"error": str(ex)[:200]}

    # ---- end to end through the C ABI with HOST (pinned) buffers ------------------------------
    # The slab of the main workload when its four host fields stay below 20 GB per rank; otherwise (N = 1 at 1024^3:
    # 35 GB of pinned memory) the 512^3 grid on a second context, and the line says so.
    e2e = None
    if not args.no_e2e:
        e_ctx, e_f, e_gd, e_dyc, e_note = ctx, f, gd, dyc, "main workload"
        if 4*gd.ncells*B > 20e9 and world == 1:
            e_gd = GridData(512, 512, 512, 25.*512, 25.*512, 25.*512, igc, 3, 1, dtype)
            e_ctx = D.Context(e_gd, local_rank)
            e_f = D.Fields(e_ctx, None)
            p1 = fill_fields_device(e_f, e_gd, seed=2, noise=0.01)
            e_ctx.set_basestate(p1["rhoref"], p1["rhorefh"], p1["thref"], p1["threfh"])
            e_dyc = D.Dycore(e_ctx, prm)
            e_note = "512x512x512 (the four host fields of the main workload would need 35 GB of pinned memory)"
        names = ["u", "v", "w", "th"]
        host = {n: torch.empty(e_gd.shape, dtype=ctx.torch_dtype).pin_memory() for n in names}
        for n in names:
            host[n].copy_(e_f[n])
        for n in ("ut", "vt", "wt", "tht"):
            e_f[n].zero_()
        k_e2e = max(1, min(args.steps, args.e2e_steps))
        e_dyc.step_host(e_f, dt, 1, host["u"], host["v"], host["w"], [host["th"]])   # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e_dyc.step_host(e_f, dt, 1, host["u"], host["v"], host["w"], [host["th"]])
        torch.cuda.synchronize()
        t_e2e = (time.perf_counter() - t0)
        if world > 1:
            t = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_e2e = float(t.item())
        nbytes = 4*e_gd.ncells*B
        e2e = {"value": world*(e_gd.npoints//world)*k_e2e/t_e2e, "unit": UNIT, "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "steps": k_e2e, "workload": e_note,
               "note": "every step copies u, v, w, th in and out over PCIe; step n+1 consumes the host result of step n, "
                       "so copies and compute cannot overlap across steps (PCIe-bound by construction)"}
        if e_ctx is not ctx:
            e_ctx.close(); del e_f; torch.cuda.empty_cache()
        del host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    ms_per_step = ms/args.steps
    value = world*npts*args.steps/(ms*1e-3)
    # dominant kernel from the live CUDA-event profile
    top = max(prof.items(), key=lambda kv: kv[1]["ms"]) if prof else (None, None)
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of the 512^3 fp64 workload
    # (profiles/r01/ncu_full_mom3_evisc_raw.csv: dram__bytes_read.sum + dram__bytes_write.sum), scaled by the point count
    ncu_traffic_per_point = {("mom3_kernel", "f64"): (10.547e9 + 4.303e9)/512**3}
    roofline = None
    if top[0] is not None:
        name, st = top
        per_launch_ms = st["ms"]/st["n"]
        passes = ALG_PASSES.get(name, 0)
        achieved = passes*npts*B/(per_launch_ms*1e-3)/1e9
        tpp = ncu_traffic_per_point.get((name, args.dtype))
        roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved/peaks["hbm_gbs"],
                    "traffic": tpp*npts if (tpp and world == 1) else None,
                    "traffic_unit": "bytes per launch (ncu capture at 512^3 under profiles/, scaled by the number of points)",
                    "peak_source": peak_src, "share_of_step": st["ms"]/ms, "algorithmic_passes": passes,
                    "ms_per_launch": per_launch_ms}
    step_alg_bytes = 3*algorithmic_bytes_per_point_substep(S, B)*npts
    whole = {"algorithmic_bytes_per_step": step_alg_bytes,
             "achieved_gbs": step_alg_bytes/(ms_per_step*1e-3)/1e9,
             "frac_of_hbm": step_alg_bytes/(ms_per_step*1e-3)/1e9/peaks["hbm_gbs"]}
    # NVLink roofline of the transposes: 2 * N_local * sizeof(TF) * (1 - 1/P) bytes leave every GPU per sub-step.
    # They are the store phases of the x-forward and the fused y-backward kernels: time = those kernels + their barriers.
    nvlink = None
    if world > 1:
        tb = 2*npts*B*(1. - 1./world)
        tms = sum(prof.get(k, {"ms": 0.})["ms"] for k in ("fft_x_forward_kernel", "tdma_fft_y_backward_kernel", "fft_y_backward_kernel",
                                                            "transpose_xy_barrier", "transpose_yx_barrier",
                                                            "all_to_all_xy_nccl", "all_to_all_yx_nccl"))/(3*args.steps)
        hms = sum(prof.get(k, {"ms": 0.})["ms"] for k in ("halo_push_kernel", "halo_barrier", "halo_unpack_kernel",
                                                            "halo_pack_kernel", "halo_sendrecv_nccl"))/(3*args.steps)
        nvlink = {"transpose_bytes_per_substep": tb, "transpose_ms_per_substep": tms,
                  "achieved_gbs": tb/(tms*1e-3)/1e9 if tms > 0 else None, "peak_gbs": 770.0,
                  "peak_source": "measured peer copy per direction (B200_PROFILING.md); nominal 900",
                  "frac": tb/(tms*1e-3)/1e9/770.0 if tms > 0 else None,
                  "note": "the transposing kernels also do the FFT work and the HBM traffic of their stage: this is a lower bound of the link rate",
                  "halo_ms_per_substep": hms}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            cpu = cpu_baseline(dtype, args.cpu_sample, dt)
        except Exception as ex:   # the baseline is a report, never a reason to lose the GPU number
            cpu = {"value": None, "unit": UNIT, "error": str(ex)[:200]}

    other = None
    if world == 1 and not args.no_side_configs:
        ctx.close(); del f; torch.cuda.empty_cache()
        other = [
            side_config(torch, D, GridData, fill_fields_device, "drycblles-shaped LES 512^3 fp64 (round-1 workload)", (512, 512, 512), np.float64, 1, peaks, args.steps),
            side_config(torch, D, GridData, fill_fields_device, "drycblles as shipped: swadvec=2 + smag2, 512^3 fp64 (Advec_2's fluxes in the fused TMA kernel)", (512, 512, 512), np.float64, 1, peaks, args.steps, swadvec="2"),
            side_config(torch, D, GridData, fill_fields_device, "drycblles 128^3 fp64 (configs[0]'s grid; launch-bound: CUDA-graph replay vs eager)", (128, 128, 128), np.float64, 1, peaks, 4*args.steps),
            side_config(torch, D, GridData, fill_fields_device, "bomex-shaped LES 512x512x256 fp32 (USESP), two scalars", (512, 512, 256), np.float32, 2, peaks, args.steps),
            side_config(torch, D, GridData, fill_fields_device, "moser180-shaped DNS 256x192x128 fp64 (advec_4m + diff_4 + pres_4, as cases/moser180 ships)", (256, 192, 128), np.float64, 1, peaks, args.steps, order=4),
            thermo_side("moist", "512x512x256", "f32", args.steps),       # bomex-shaped with Thermo_moist (thl + qt, base-state update on the device)
            thermo_side("buoy", "256x192x128", "f64", args.steps),        # 4th-order DNS with slope-enabled Thermo_buoy (drycblslope-like)
        ]

    itot, jtot, ktot_l = gd.imax, gd.jmax, gd.kmax
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic (generated on the device)",
            "config": {"workload": f"drycblles-shaped LES, global grid {itot_g}x{jtot_g}x{ktot} ({itot}x{jtot}x{ktot_l} points per GPU), advec_{args.swadvec}+diff_smag2+pres_2+thermo_dry, S=1",
                       "global_grid": f"{itot_g}x{jtot_g}x{ktot}",
                       "parallelism": (f"one domain in {world} y-slabs (npx=1, npy={world}); transposes and ghost rows: "
                                       + ("stores of the FFT / pack kernels straight into peer memory over NVLink (CUDA IPC), 4-byte NCCL all-reduce as barrier"
                                          if ctx.transport == "peer" else "grouped ncclSend/ncclRecv")) if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (each field >> 126 MB)", "dt": dt},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "whole_step_roofline": whole, "nvlink": nvlink, "cpu_baseline": cpu,
            "kernels_ms_per_step": {k: v["ms"]/args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
            "finite": finite, "post_step_divergence": post_div, "other_configs": other}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
def cpu_baseline(dtype, sample, dt, max_threads=None, min_seconds=8.0):
    """Reference CPU kernels (oracle/_ref, built from the reference's own sources with its release
    flags) -- or the numpy port when that library is absent -- timed on the host cores: every
    thread advances its own independent sub-domain of shape `sample` by full RK3 steps."""
    from oracle import oracle as O, step as ostep, refbind
    from microhh_b200.grid import GridData
    from microhh_b200.synthetic import make_case
    it, jt, kt = parse_workload(sample)
    cores = max_threads or (os.cpu_count() or 1)
    use_ref = refbind.available(fast=True)
    kind = "reference" if use_ref else "port"
    if not use_ref:
        cores = min(cores, 4)
    work = []
    for t in range(cores):
        g = O.Grid(it, jt, kt, 25.*it, 25.*jt, 25.*kt, 3, 3, 1, dtype)
        gd = GridData(it, jt, kt, 25.*it, 25.*jt, 25.*kt, 3, 3, 1, dtype)
        K = refbind.RefKernels(g, fast=True) if use_ref else O.NumpyKernels(g)
        work.append((g, K, make_case(gd, seed=10 + t, noise=0.01)))
    prm = ostep.default_params()
    counts = [0]*cores
    stop = threading.Event()

    def worker(i):
        g, K, c = work[i]
        pres = None
        while not stop.is_set():
            for ss in range(3):
                pres = ostep.dycore_substep(g, K, c, prm, ss, dt, pres)
            counts[i] += 1

    # NOTE: ref_set_geom is process-global in the harness; all threads use the same geometry.
    threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(cores)]
    t0 = time.perf_counter()
    for th in threads:
        th.start()
    time.sleep(min_seconds)
    stop.set()
    for th in threads:
        th.join()
    el = time.perf_counter() - t0
    total = sum(counts)
    return {"value": total*it*jt*kt/el, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{total} RK3 steps of {cores} independent {it}x{jt}x{kt} sub-domains (one per thread) in {el:.1f} s; "
                      "FFT = numpy pocketfft stand-in for FFTW"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dtype = np.float64 if args.dtype == "f64" else np.float32
    itot, jtot, ktot = parse_workload(args.workload)
    t0 = time.perf_counter()
    secs = max(4.0, min(20.0, 4.0*(args.steps + args.warmup)))
    cpu = cpu_baseline(dtype, args.cpu_sample, args.dt, min_seconds=secs)
    el = time.perf_counter() - t0
    it, jt, kt = parse_workload(args.cpu_sample)
    line = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT,
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3*it*jt*kt*cpu["cores"]/cpu["value"] if cpu["value"] else None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"drycblles-shaped LES, global grid {itot}x{jtot}x{ktot}, advec_2i5+diff_smag2+pres_2+thermo_dry, S=1",
                       "note": "CPU arm times a bounded sample of the same workload: " + cpu["sample"]},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": el}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("MHH_BENCH_WORKLOAD", "1024x1024x1024"),
                    help="the GLOBAL grid (strong scaling, default) or the per-GPU grid (--scaling weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--no-side-configs", action="store_true")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--dt", type=float, default=1.0)
    ap.add_argument("--swadvec", default="2i5", choices=["2i5", "2"], help="2 = cases/drycblles as shipped (side measurement; the headline is 2i5)")
    ap.add_argument("--igc", type=int, default=0, help="x ghost cells (0 = 4: what the adapters request; 3 = the reference's minimum)")
    ap.add_argument("--cpu-sample", default="64x64x64")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
