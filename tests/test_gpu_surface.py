"""
Monin-Obukhov surface model on the device: `mhh_boundary_surface_exec` against the reference's own compiled
kernels (oracle/_ref: stability, surfm, surfs, calc_dutot, calc_duvdz_mo, calc_dbdz_mo; numpy restatement when the library is
absent), and a SELF-DRIVEN multi-step LES -- Model::exec's order with the surface model inside every sub-step
(src/model.cxx:368-504) -- against the oracle stepping the same way.
"""
import copy
import numpy as np
import pytest

from util import TOL, rel_l2, make_pair, interior, prepare_halos
from oracle import oracle as O
from oracle import step as ostep
from oracle import refbind

pytestmark = pytest.mark.gpu


def oracle_surface(g, z0m, z0h, thermobc):
    if refbind.available(False):
        return refbind.RefSurface(g, z0m, z0h, O.BC_DIRICHLET, thermobc)
    return O.BoundarySurface(g, z0m, z0h, O.BC_DIRICHLET, thermobc)


def kernels(g):
    return refbind.RefKernels(g, fast=False) if refbind.available(False) else O.NumpyKernels(g)


def setup(shape, dtype, thermobc):
    g, gd, case = make_pair(*shape, dtype, stretched=True, anelastic=False)
    for n in ("u", "v", "th"):
        case[n + "_bot"] = np.zeros(gd.shape2d, dtype)
    case["th_bot"][...] = 300.4 if thermobc == O.BC_DIRICHLET else 0.
    from microhh_b200 import dycore as D
    ctx = D.Context(gd, 0)
    ctx.set_basestate(case["rhoref"], case["rhorefh"], case["thref"], case["threfh"])
    return g, gd, case, D, ctx


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("thermobc", [O.BC_FLUX, O.BC_DIRICHLET])
@pytest.mark.parametrize("shape", [(32, 24, 8), (96, 40, 12)])
def test_boundary_surface_exec(dtype, thermobc, shape):
    g, gd, case, D, ctx = setup(shape, dtype, thermobc)
    prepare_halos(g, case)
    f = D.Fields(ctx, case)
    prm = D.make_params()
    S = D.Boundary_surface(ctx, f, z0m=0.1, z0h=0.01, thermobc=thermobc)
    R = oracle_surface(g, 0.1, 0.01, thermobc)
    # The Obukhov length comes out of a FLOAT table searched with a float Richardson number (include/boundary_surface_kernels.h:
    # 245-290): a one-ulp difference of the double input flips the float, so everything downstream agrees to float accuracy only.
    tol = 5e-6 if dtype == np.float64 else 50*TOL[dtype]
    for it in range(2):              # the second call starts the table search from the first one's index
        S.exec(f, prm)
        R.exec(case, case["thref"], case["threfh"])
    ctx.sync()
    zsl = float(g.z[g.kstart])
    # z/L rather than L: L passes through +-infinity at neutral points
    assert rel_l2(zsl/S.obuk.cpu().numpy(), zsl/R.obuk) <= tol and rel_l2(S.ustar.cpu().numpy(), R.ustar) <= tol
    sl = (slice(g.jstart, g.jend), slice(g.istart, g.iend))
    for n in ("u_fluxbot", "v_fluxbot", "u_gradbot", "v_gradbot", "th_bot", "th_gradbot", "th_fluxbot"):
        assert rel_l2(f[n].cpu().numpy(), case[n]) <= tol, n
    for n in ("dudz_mo", "dvdz_mo", "dbdz_mo"):
        assert rel_l2(f[n].cpu().numpy()[sl], case[n][sl]) <= tol, n


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_boundary_surface_neutral(dtype):
    """Thermo_type::Disabled: stability_neutral (src/boundary_surface.cxx:136-180)."""
    g, gd, case, D, ctx = setup((32, 24, 8), dtype, O.BC_FLUX)
    prepare_halos(g, case)
    f = D.Fields(ctx, case)
    prm = D.make_params(swthermo=None)
    S = D.Boundary_surface(ctx, f, z0m=0.1, z0h=0.1, thermobc=O.BC_FLUX, sbcbot=[1])
    R = oracle_surface(g, 0.1, 0.1, O.BC_FLUX)
    S.exec(f, prm)
    case2 = copy.deepcopy(case); case2["th_bcbot"] = 1
    R.exec(case2, neutral=True)
    ctx.sync()
    tol = 50*TOL[dtype]
    assert rel_l2(S.ustar.cpu().numpy(), R.ustar) <= tol
    sl = (slice(g.jstart, g.jend), slice(g.istart, g.iend))
    for n in ("u_fluxbot", "v_fluxbot", "u_gradbot", "v_gradbot"):
        assert rel_l2(f[n].cpu().numpy(), case2[n]) <= tol, n
    for n in ("dudz_mo", "dvdz_mo"):
        assert rel_l2(f[n].cpu().numpy()[sl], case2[n][sl]) <= tol, n


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_self_driven_les_three_steps(dtype):
    """drycblles-type run with everything on the device: three full RK3 steps, the surface model recomputed in every sub-step
    (no-slip bottom, prescribed th flux), against the oracle driven the same way."""
    shape = (64, 32, 16)
    g, gd, case, D, ctx = setup(shape, dtype, O.BC_FLUX)
    f = D.Fields(ctx, case)
    prm = D.make_params(mbcbot=0)          # no-slip bottom: ghost cells from u_bot (Dirichlet), as drycblles.ini
    oprm = ostep.default_params(); oprm.update(mbcbot=O.BC_DIRICHLET)
    S = D.Boundary_surface(ctx, f, z0m=0.1, z0h=0.1, thermobc=O.BC_FLUX)
    R = oracle_surface(g, 0.1, 0.1, O.BC_FLUX)
    dyc = D.Dycore(ctx, prm)
    K = kernels(g)
    dt = 2.0
    for _ in range(3):
        dyc.step_surface(f, S, dt)
        ostep.dycore_step(g, K, case, oprm, dt, surface_model=R)
    ctx.sync()
    for n in ("u", "v", "w", "th"):
        assert rel_l2(interior(g, f[n].cpu().numpy()), interior(g, case[n])) <= 20*TOL[dtype], n
    assert rel_l2(S.ustar.cpu().numpy(), R.ustar) <= (5e-6 if dtype == np.float64 else 100*TOL[dtype])      # float lookup table
