"""Shared definitions of the golden vectors of the rows added in the last session of round 2 (Thermo_moist, Thermo_buoy,
Advec_2i4, Advec_2i62): builds the seeded inputs and runs one RK3 step with a given kernel set (the reference's compiled kernels
in tests/make_golden_new_rows.py, the numpy oracle in tests/test_golden_new_rows.py)."""
import hashlib

import numpy as np

from util import make_moist_pair, stretched_z
from oracle import oracle as O, step as ostep

CASES = {
    # name: (kind, dtype, options)
    "moist_smag2_update_16x12x20_f64": ("moist", np.float64, dict(swdiff="smag2", update=True, cold=False)),
    "moist_smag2_update_16x12x20_f32": ("moist", np.float32, dict(swdiff="smag2", update=True, cold=False)),
    "moist_diff2_cold_16x12x20_f64": ("moist", np.float64, dict(swdiff="2", update=True, cold=True)),
    "buoy_o4_slope_baroclinic_16x12x10_f64": ("buoy4", np.float64, dict(tb=dict(alpha=0.2, n2=1.e-2, utrans=0.1, swbaroclinic=True, dbdy_ls=2.e-3), swadvec="4m")),
    "buoy_o4_plain_16x12x10_f32": ("buoy4", np.float32, dict(tb={}, swadvec="4")),
    "advec_2i4_smag2_dry_16x12x10_f64": ("2ix", np.float64, dict(scheme="2i4", gc=(2, 2, 2), swdiff="smag2", lim=())),
    "advec_2i62_limited_16x12x10_f32": ("2ix", np.float32, dict(scheme="2i62", gc=(3, 3, 2), swdiff="2", lim=("s1",))),
}
NAMES = {"moist": ("u", "v", "w", "thl", "qt", "p"), "buoy4": ("u", "v", "w", "th", "p"), "2ix": ("u", "v", "w", "th", "s1", "p")}


def build(name):
    """-> (g, case, prm, dt, pres_order) for a golden case"""
    from microhh_b200.grid import GridData
    from microhh_b200.synthetic import make_case
    kind, dtype, o = CASES[name]
    if kind == "moist":
        g, gd, case, pbot = make_moist_pair(16, 12, 20, dtype, cold=o["cold"])
        smag = o["swdiff"] == "smag2"
        prm = ostep.default_params()
        prm.update(swdiff=o["swdiff"], swthermo="moist", thermo_moist=dict(pbot=pbot, swupdatebasestate=o["update"]), surface_model=smag,
                   visc=1e-5 if smag else 1e-2, svisc=1e-5 if smag else 1e-2)
        return g, case, prm, 2.0, 2
    if kind == "buoy4":
        shape = (16, 12, 10)
        z = stretched_z(shape[2], 2.)
        g = O.Grid(*shape, 2*np.pi, np.pi, 2., 3, 3, 3, dtype, z=z, order=4)
        gd = GridData(*shape, 2*np.pi, np.pi, 2., 3, 3, 3, dtype, z=z, order=4)
        case = make_case(gd, seed=5, noise=0.02)
        case["w"][:g.kstart+1] = 0; case["w"][g.kend:] = 0
        for n in ("u", "v"):
            for sfx in ("_bot", "_top", "_gradbot", "_gradtop"):
                case[n + sfx] = np.zeros(gd.shape2d, dtype)
        case["th_gradbot"] = np.full(gd.shape2d, -0.3, dtype); case["th_gradtop"] = np.full(gd.shape2d, 0.2, dtype)
        case["th"] = (case["th"] - dtype(300.)).astype(dtype)          # scalar 0 is the buoyancy
        prm = ostep.default_params()
        prm.update(swadvec=o["swadvec"], swdiff="4", swthermo="buoy", thermo_buoy=o["tb"], visc=1e-3, svisc=1e-3, mbcbot=0, mbctop=0)
        return g, case, prm, 1e-3, 4
    shape = (16, 12, 10)
    z = stretched_z(shape[2], 3200.)
    g = O.Grid(*shape, 3200., 3200., 3200., *o["gc"], dtype, z=z)
    gd = GridData(*shape, 3200., 3200., 3200., *o["gc"], dtype, z=z)
    case = make_case(gd, seed=4, anelastic=True, ns=2)
    smag = o["swdiff"] == "smag2"
    prm = ostep.default_params()
    prm.update(swadvec=o["scheme"], swdiff=o["swdiff"], surface_model=smag, visc=1e-5 if smag else 1e-2, svisc=1e-5 if smag else 1e-2,
               fluxlimit_list=o["lim"])
    return g, case, prm, 2.0, 2


def digest(case, names):
    h = hashlib.sha256()
    for n in names:
        if n != "p":
            h.update(np.ascontiguousarray(case[n]).tobytes())
    return h.hexdigest()
