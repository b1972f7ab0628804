"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import numpy as np

from oracle import oracle as O
from oracle import step as ostep
from microhh_b200.grid import GridData
from microhh_b200.synthetic import make_case

# tolerances from BASELINE.json north_star: relative L2 <= 1e-12 (fp64), <= 1e-5 (fp32)
TOL = {np.float64: 1e-12, np.float32: 1e-5}


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    den = np.sqrt((b*b).sum())
    return np.sqrt(((a-b)**2).sum()) / (den if den > 0 else 1.)


def stretched_z(ktot, zsize):
    dz = np.linspace(0.6, 1.4, ktot); dz *= zsize/dz.sum()
    zh = np.concatenate([[0.], np.cumsum(dz)])
    return 0.5*(zh[1:] + zh[:-1])


def make_pair(itot, jtot, ktot, dtype, stretched=False, anelastic=False, ns=1, seed=2, sizes=(3200., 3200., 3200.), igc=3):
    """igc = 4 is what the USESP adapters request (Grid::set_minimum_ghost_cells): the fp32 row pitch is then a multiple
    of 16 bytes and the TMA-staged kernels apply."""
    z = stretched_z(ktot, sizes[2]) if stretched else None
    g = O.Grid(itot, jtot, ktot, *sizes, igc, 3, 1, dtype, z=z)
    gd = GridData(itot, jtot, ktot, *sizes, igc, 3, 1, dtype, z=z)
    case = make_case(gd, seed=seed, anelastic=anelastic, ns=ns)
    return g, gd, case


def interior(g, a, k0=None, k1=None):
    k0 = g.kstart if k0 is None else k0
    k1 = g.kend if k1 is None else k1
    return a[k0:k1, g.jstart:g.jend, g.istart:g.iend]


def prepare_halos(g, case):
    """cyclic + vertical ghost cells with the oracle so that single-kernel tests see valid halos."""
    prm = ostep.default_params()
    for n in ["u", "v", "w"] + case["scalars"]:
        O.boundary_cyclic(g, case[n])
    for n in ["u", "v"]:
        O.ghost_cells_bot_2nd(g, case[n], prm["mbcbot"], None, case[n + "_gradbot"])
        O.ghost_cells_top_2nd(g, case[n], prm["mbctop"], None, case[n + "_gradtop"])
    for s in case["scalars"]:
        O.ghost_cells_bot_2nd(g, case[s], prm["sbcbot"], None, case[s + "_gradbot"])
        O.ghost_cells_top_2nd(g, case[s], prm["sbctop"], None, case[s + "_gradtop"])


def add_sgstke(g, case, seed=5, stable_bottom=True):
    """Extends a synthetic case with the prognostic SGS TKE of Diff_tke2 (src/diff_tke2.cxx:544): a positive field with a few
    values below Constants::sgstke_min, its tendency, zero-flux 2-D companions and the `eviscs` diagnostic.  With
    `stable_bottom` half of the surface points get a positive db/dz so that both branches of the length scale run."""
    TF = g.TF
    rng = np.random.default_rng(seed)
    shape = case["u"].shape
    e = 0.05 + 0.4*rng.random(shape)
    e[rng.random(shape) < 0.02] = 1.e-9
    case["sgstke"] = np.ascontiguousarray(e.astype(TF))
    case["sgstket"] = np.zeros(shape, TF)
    case["eviscs"] = np.zeros(shape, TF)
    for n in ("fluxbot", "fluxtop", "gradbot", "gradtop"):
        case["sgstke_" + n] = np.zeros(shape[1:], TF)
    if "sgstke" not in case["scalars"]:
        case["scalars"] = list(case["scalars"]) + ["sgstke"]
    if stable_bottom:
        d = case["dbdz_mo"].copy()
        d[:, ::2] = -d[:, ::2]
        case["dbdz_mo"] = d
    return case


def moist_case(g, gd, dtype, seed=3, cold=False):
    """A bomex-like (or, `cold`, mixed-phase) moist state on the test grid: thl rising with height, qt falling, noise, and a
    moist layer that saturates a good share of the points so that the Newton loops of sat_adjust run."""
    rng = np.random.default_rng(seed)
    zfull = np.asarray(g.z, np.float64)[:, None, None]
    zrel = zfull/float(g.zsize)
    t0 = 262. if cold else 298.
    thl = t0 + 6.*zrel + 0.3*rng.standard_normal(gd.shape)
    qsurf = 2.4e-3 if cold else 17.e-3
    qt = qsurf*(1. - 0.5*zrel) + (0.8e-3 if cold else 3.e-3)*np.exp(-((zrel - 0.45)/0.15)**2) + 1.e-4*rng.standard_normal(gd.shape)
    return thl.astype(dtype), np.maximum(qt, 1e-5).astype(dtype)


def make_moist_pair(itot, jtot, ktot, dtype, cold=False, igc=3, sizes=(3200., 3200., 3000.), anelastic=True):
    """Grid pair + a synthetic moist case: scalars thl and qt (moist_case), zero surface fluxes replaced by bomex-like ones, and
    the initial base state `moist_bs` from the mean profiles (create_basestate: calc_top_and_bot + calc_base_state)."""
    g, gd, case = make_pair(itot, jtot, ktot, dtype, stretched=True, anelastic=anelastic, ns=2, sizes=sizes, igc=igc)
    thl, qt = moist_case(g, gd, dtype, cold=cold)
    ren = {"th": "thl", "s1": "qt"}
    for old, new in ren.items():
        for k in [k for k in case if k == old or k.startswith(old + "_") or k == old + "t"]:
            case[new + k[len(old):]] = case.pop(k)
    case["scalars"] = ["thl", "qt"]
    case["thl"], case["qt"] = thl, qt
    case["thl_fluxbot"] = np.full(gd.shape2d, 8.e-3, dtype); case["qt_fluxbot"] = np.full(gd.shape2d, 5.2e-5, dtype)
    case["thl_gradbot"] = np.full(gd.shape2d, -1.e-3, dtype); case["qt_gradbot"] = np.full(gd.shape2d, -1.e-6, dtype)
    case["thl_gradtop"] = np.full(gd.shape2d, 3.e-3, dtype); case["qt_gradtop"] = np.full(gd.shape2d, -1.e-6, dtype)
    pbot = 70000. if cold else 101500.
    thl0, qt0 = O.mean_profile(g, thl), O.mean_profile(g, qt)
    O.moist_top_and_bot(g, thl0, qt0)
    case["moist_bs"] = O.moist_base_state(g, thl0, qt0, pbot)
    case["moist_ref"] = (thl0, qt0)
    if anelastic:
        case["rhoref"] = case["moist_bs"]["rhoref"].copy(); case["rhorefh"] = case["moist_bs"]["rhorefh"].copy()
        # the dynamics read rhoref at the ghost levels too (advection weights): extend like Fields does (constant extrapolation)
        for a, lo, hi in ((case["rhoref"], g.kstart, g.kend - 1), (case["rhorefh"], g.kstart, g.kend)):
            a[:lo] = a[lo]; a[hi+1:] = a[hi]
    return g, gd, case, pbot
