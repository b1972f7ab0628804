"""Side measurement of the thermo couplings added after the headline bench: one RK3 step with
  --thermo moist : LES (2i5 + smag2), prognostic thl + qt, Thermo_moist registered into the fused sub-step with
                   swupdatebasestate (mean profiles + base state on the device, buoyancy through the saturation adjustment)
  --thermo buoy  : 4th-order DNS (4m + 4 + pres_4), prognostic b, slope-enabled Thermo_buoy registered into the fused sub-step
per-kernel CUDA-event times of the thermo kernels beside the step.  python tools/thermo_bench.py [--thermo moist]
[--grid 512x512x256] [--dtype f32] [--steps 5]"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(thermo="moist", grid="512x512x256", dtype="f32", steps=5):
    """One measurement; returns the JSON-able dict (also used by bench.py for its thermo side lines)."""
    import types
    a = types.SimpleNamespace(thermo=thermo, grid=grid, dtype=dtype, steps=steps)
    import torch
    from microhh_b200 import dycore as D
    from microhh_b200.grid import GridData
    from microhh_b200.synthetic import fill_fields_device
    it, jt, kt = (int(x) for x in a.grid.split("x"))
    dtype = np.float64 if a.dtype == "f64" else np.float32
    B = np.dtype(dtype).itemsize
    npts = it*jt*kt
    moist = a.thermo == "moist"
    if moist:
        gd = GridData(it, jt, kt, 25.*it, 25.*jt, 3000., 4, 3, 1, dtype)
        scal = ["thl", "qt"]
    else:
        gd = GridData(it, jt, kt, 6.28, 3.14, 2., 3, 3, 3, dtype, order=4)
        scal = ["b"]
    ctx = D.Context(gd, 0)
    f = D.Fields(ctx, None, scalars=scal, visc=1e-5 if moist else 1e-3, svisc=1e-5 if moist else 1e-3)
    # fill_fields_device writes u, v, w and a scalar called th: borrow the first scalar's tensor under that name
    f.t["th"] = f.t[scal[0]]
    prof1d = fill_fields_device(f, gd, noise=0.01)
    del f.t["th"]
    ks, ke = gd.kstart, gd.kend
    if moist:
        z = torch.from_numpy(np.asarray(gd.z, np.float64)).to(f["u"].device)
        zrel = (z/float(gd.zsize))[:, None, None]
        f["thl"].copy_((298. + 6.*zrel + (f["thl"].double() - 300. - 0.003*z[:, None, None])).to(f["thl"].dtype))
        # a cumulus-like layer: the mean profile stays just below saturation and the fluctuations saturate ~1 % of all points, up to
        # ~17 % in the core of the layer (tuned on the CPU with the oracle's saturation adjustment on a sample of this very field;
        # with 1.1e-3 the field of the first measurements was cloud-free)
        f["qt"].copy_((17.e-3*(1. - 0.75*zrel) + 1.5e-3*torch.exp(-((zrel - 0.45)/0.15)**2)
                       + 1.e-4*torch.randn(gd.shape, device=z.device, dtype=torch.float64)).clamp_min(1e-5).to(f["qt"].dtype))
        for n, v in (("thl_fluxbot", 8.e-3), ("qt_fluxbot", 5.2e-5), ("thl_gradbot", -1.e-3), ("qt_gradbot", -1.e-6), ("qt_gradtop", -1.e-6)):
            f[n].fill_(v)
        ctx.set_basestate(prof1d["rhoref"], prof1d["rhorefh"], prof1d["thref"], prof1d["threfh"])
        means = []
        for n in scal:
            m = f[n][:, gd.jstart:gd.jend, gd.istart:gd.iend].double().mean(dim=(1, 2)).cpu().numpy()
            m[:ks] = m[ks]; m[ke:] = m[ke-1]
            means.append(m.astype(dtype))
        T = D.Thermo_moist(ctx, f, 101500., swupdatebasestate=True)
        T.calc_base_state(*means)
        prm = D.make_params(swadvec="2i5", swdiff="smag2", swthermo="moist", ns=2)
        names = ("moist_mean_profile_kernel", "moist_base_state_kernel", "moist_buoyancy_tend_kernel")
        passes = {"moist_mean_profile_kernel": 2, "moist_buoyancy_tend_kernel": 4}
    else:
        f["b"].copy_((0.05*(f["b"].double() - 300.)).to(f["b"].dtype))
        ones = np.ones(gd.kcells, dtype)
        ctx.set_basestate(ones, ones, 300*ones, 300*ones)
        T = D.Thermo_buoy(ctx, alpha=0.1, n2=3., utrans=0.)
        prm = D.make_params(swadvec="4m", swdiff="4", swthermo="buoy", surface_model=False, mbcbot=0, mbctop=0)
        names = ("thermo_buoy_kernel",)
        passes = {"thermo_buoy_kernel": 9}            # slope-enabled: R b, u, w (3) + RMW ut, wt, bt (6)
    T.register()
    dyc = D.Dycore(ctx, prm)
    dt = 1.0 if moist else 1e-3
    for _ in range(3):
        dyc.step(f, dt)
    torch.cuda.synchronize()
    ctx.profile_start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        dyc.step(f, dt)
    e1.record(); torch.cuda.synchronize()
    prof = ctx.profile_stop()
    ms = e0.elapsed_time(e1)/a.steps
    kern = {}
    for n in names:
        k = prof.get(n, {"ms": 0., "n": 0})
        per = k["ms"]/max(k["n"], 1)
        kern[n] = {"launches_per_step": k["n"]/a.steps, "ms_per_launch": per, "ms_per_step": k["ms"]/a.steps}
        if n in passes and per > 0:
            kern[n].update(algorithmic_passes=passes[n], achieved_gbs=passes[n]*npts*B/(per*1e-3)/1e9)
    out = {"workload": f"{'LES 2i5+smag2, thl+qt, Thermo_moist (swupdatebasestate)' if moist else '4th-order DNS 4m+4+pres_4, Thermo_buoy (slope)'}, {a.grid} {a.dtype}",
           "ms_per_step": ms, "value": npts/(ms*1e-3), "unit": "grid-point-steps/s",
           "finite": bool(torch.isfinite(f["u"]).all().item() and torch.isfinite(f["w"]).all().item()),
           "thermo_share_of_step": sum(v["ms_per_step"] for v in kern.values())/ms,
           "thermo_kernels": kern,
           "kernels_ms_per_step": {n: v["ms"]/a.steps for n, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]}}
    if moist:
        out["nonconverged"] = T.nonconverged()
        out["base_state_sweeps_last"] = T.base_state_sweeps()
        T.get_thermo_field(f["evisc"], "ql", f)
        out["cloud_fraction"] = float((f["evisc"][ks:ke, gd.jstart:gd.jend, gd.istart:gd.iend] > 0).double().mean().item())
        bs = T.get_profiles()
        out["thvrefh_surface"] = float(bs["thvrefh"][ks]); out["prefh_top"] = float(bs["prefh"][ke])
    T.unregister()
    ctx.close()
    del f
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--thermo", default="moist"); ap.add_argument("--grid", default="512x512x256")
    ap.add_argument("--dtype", default="f32"); ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    print(json.dumps(run(a.thermo, a.grid, a.dtype, a.steps)))


if __name__ == "__main__":
    main()
