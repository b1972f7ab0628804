"""Side measurement: one RK3 step of the LES path with the Deardorff SGS-TKE closure (swdiff = tke2: th + sgstke prognostic,
exec_viscosity fused into tke2_visc_kernel) on a synthetic grid, per-kernel CUDA-event times.  python tools/tke2_bench.py
[--grid 256x256x256] [--dtype f64] [--steps 5]"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="256x256x256"); ap.add_argument("--dtype", default="f64"); ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import torch
    from microhh_b200 import dycore as D
    from microhh_b200.grid import GridData
    from microhh_b200.synthetic import fill_fields_device
    it, jt, kt = (int(x) for x in a.grid.split("x"))
    dtype = np.float64 if a.dtype == "f64" else np.float32
    gd = GridData(it, jt, kt, 25.*it, 25.*jt, 25.*kt, 4, 3, 1, dtype)
    ctx = D.Context(gd, 0)
    f = D.Fields(ctx, None, scalars=["th", "sgstke"])
    prof1d = fill_fields_device(f, gd, noise=0.01)
    ctx.set_basestate(prof1d["rhoref"], prof1d["rhorefh"], prof1d["thref"], prof1d["threfh"])
    f["sgstke"].fill_(0.3)
    f["sgstke"].mul_(1. + 0.5*torch.rand_like(f["sgstke"]))
    prm = D.make_params(swdiff="tke2", ns=2)
    T = D.Diff_tke2(ctx, prm, f); T.create(f); T.register()
    dyc = D.Dycore(ctx, prm)
    for _ in range(3):
        dyc.step(f, 1.0)
    torch.cuda.synchronize()
    ctx.profile_start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        dyc.step(f, 1.0)
    e1.record(); torch.cuda.synchronize()
    prof = ctx.profile_stop()
    ms = e0.elapsed_time(e1)/a.steps
    npts = it*jt*kt; B = np.dtype(dtype).itemsize
    k = prof.get("tke2_visc_kernel", {"ms": 0., "n": 1})
    per = k["ms"]/max(k["n"], 1)
    out = {"workload": f"LES with swdiff=tke2 (th + sgstke), {a.grid} {a.dtype}", "ms_per_step": ms,
           "value": npts/(ms*1e-3), "unit": "grid-point-steps/s", "finite": bool(torch.isfinite(f["u"]).all().item()),
           "sgstke_min": float(f["sgstke"][gd.kstart:gd.kend, gd.jstart:gd.jend, gd.istart:gd.iend].min().item()),
           "tke2_visc_kernel": {"ms_per_launch": per, "algorithmic_passes": 9, "achieved_gbs": 9*npts*B/(per*1e-3)/1e9 if per > 0 else None},
           "kernels_ms_per_step": {n: v["ms"]/a.steps for n, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:10]}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
