#!/usr/bin/env python
"""Times one full RK3 step of the 4th-order DNS configuration (advec_4 + diff_4 + pres_4) on a moser180-shaped grid
(256 x 192 x 128, fp64): context line for DESIGN.md, not the headline bench."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from microhh_b200 import dycore as D
from microhh_b200.grid import GridData
from microhh_b200.synthetic import make_case


def main():
    it, jt, kt = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "256x192x128").split("x"))
    dtype = np.float64
    dz = np.linspace(0.6, 1.4, kt); dz *= 2./dz.sum(); zh = np.concatenate([[0.], np.cumsum(dz)])
    gd = GridData(it, jt, kt, 2*np.pi, np.pi, 2., 3, 3, 3, dtype, z=0.5*(zh[1:] + zh[:-1]), order=4)
    case = make_case(gd, seed=5, noise=0.02)
    case["w"][:gd.kstart+1] = 0; case["w"][gd.kend:] = 0
    case["th"] = (1. + 0.1*case["u"]).astype(dtype)
    for n in ("u", "v"):
        for sfx in ("_bot", "_top", "_gradbot", "_gradtop"):
            case[n + sfx] = np.zeros(gd.shape2d, dtype)
    case["th_gradbot"] = np.zeros(gd.shape2d, dtype); case["th_gradtop"] = np.zeros(gd.shape2d, dtype)
    ctx = D.Context(gd, 0)
    ones = np.ones(gd.kcells, dtype)
    ctx.set_basestate(ones, ones, 300*ones, 300*ones)
    f = D.Fields(ctx, case, visc=1e-4, svisc=1e-4)
    prm = D.make_params(swadvec="4", swdiff="4", swthermo=None, surface_model=False, mbcbot=0, mbctop=0)
    dyc = D.Dycore(ctx, prm)
    dt = 1e-4
    for _ in range(3):
        dyc.step(f, dt)
    torch.cuda.synchronize()
    ctx.profile_start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        dyc.step(f, dt)
    e1.record(); torch.cuda.synchronize()
    prof = ctx.profile_stop()
    ms = e0.elapsed_time(e1)/n
    print(json.dumps({"config": f"4th-order DNS {it}x{jt}x{kt} fp64 (advec_4+diff_4+pres_4), S=1", "ms_per_step": ms,
                      "grid_point_steps_per_s": it*jt*kt/(ms*1e-3), "finite": bool(torch.isfinite(f["u"]).all().item()),
                      "kernels_ms_per_step": {k: v["ms"]/n for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}}))


if __name__ == "__main__":
    main()
