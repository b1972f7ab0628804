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
