"""
Thermo_buoy on the device (reference src/thermo_buoy.cxx): the single fused exec kernel, get_thermo_field("N2") and the fused
sub-steps with swthermo = buoy (4th-order DNS cases: drycbl, drycblslope, prandtlslope, rayleighbenard ...; 2nd order with
diff_2), through the C ABI against the oracle (pinned bit for bit to the compiled reference in tests/test_oracle_vs_ref.py).
Tolerances: relative L2 <= 1e-12 (fp64), <= 1e-5 (fp32).
"""
import copy
import numpy as np
import pytest

from util import TOL, rel_l2, make_pair, interior, stretched_z
from oracle import oracle as O
from oracle import step as ostep

pytestmark = pytest.mark.gpu

DTYPES = [np.float64, np.float32]
VARIANTS = [dict(),                                                                  # plain: wt += interp(b)
            dict(alpha=0.17, n2=2.e-3, utrans=0.4),                                   # slope-enabled
            dict(n2=3.e-3),                                                           # has_N2 alone switches the slope kernels on
            dict(swbaroclinic=True, dbdy_ls=3.e-4),                                   # baroclinic only
            dict(alpha=0.5235, n2=1., swbaroclinic=True, dbdy_ls=2.e-3)]              # everything (alpha of cases/prandtlslope)


def grids(order, shape, dtype, stretched=True):
    from microhh_b200.grid import GridData
    it, jt, kt = shape
    if order == 4:
        z = stretched_z(kt, 2.) if stretched else None
        g = O.Grid(it, jt, kt, 6., 4., 2., 3, 3, 3, dtype, z=z, order=4)
        gd = GridData(it, jt, kt, 6., 4., 2., 3, 3, 3, dtype, z=z, order=4)
        return g, gd
    g, gd, _ = make_pair(it, jt, kt, dtype, stretched=stretched)
    return g, gd


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("shape", [(32, 16, 12), (60, 9, 10), (24, 1, 8)])
def test_thermo_buoy_exec(dtype, order, shape):
    from microhh_b200 import dycore as D
    g, gd = grids(order, shape, dtype)
    rng = np.random.default_rng(11)
    fld = lambda: rng.standard_normal(gd.shape).astype(dtype)
    case = dict(u=fld(), v=fld(), w=fld(), th=(0.1*fld()).astype(dtype), ut=fld(), vt=fld(), wt=fld(), tht=fld())
    ctx = D.Context(gd, 0)
    ones = np.ones(gd.kcells, dtype)
    ctx.set_basestate(ones, ones, 300*ones, 300*ones)
    K = O.NumpyKernels(g)
    for tb in VARIANTS:
        f = D.Fields(ctx, case, visc=0.7, svisc=1.3)
        D.Thermo_buoy(ctx, **tb).exec(f)
        ref = {n: case[n].copy() for n in ("ut", "vt", "wt", "tht", "u", "v", "w", "th")}
        O.thermo_buoy_exec(K, dict(ref, scalars=["th"]), tb, order)
        for n in ("ut", "vt", "wt", "tht"):
            got = f[n].cpu().numpy()
            assert rel_l2(got, ref[n]) <= TOL[dtype], (tb, n)
            assert np.array_equal(got[:g.kstart], case[n][:g.kstart]) and np.array_equal(got[g.kend:], case[n][g.kend:]), (tb, n)   # ghost levels untouched
        assert np.array_equal(f["vt"].cpu().numpy(), case["vt"])                                  # no v tendency in any variant
        assert np.array_equal(f["wt"].cpu().numpy()[g.kstart], case["wt"][g.kstart])              # the wall level of wt stays
        if not tb:
            assert np.array_equal(f["ut"].cpu().numpy(), case["ut"]) and np.array_equal(f["tht"].cpu().numpy(), case["tht"])
        else:
            assert not np.array_equal(f["tht"].cpu().numpy(), case["tht"])
    # get_thermo_field("N2")
    f = D.Fields(ctx, case, visc=0.7, svisc=1.3)
    T = D.Thermo_buoy(ctx, n2=1.5e-4)
    T.get_thermo_field_N2(f["evisc"], f)
    n2 = np.zeros(gd.shape, dtype)
    K.thermo_buoy_N2(n2, case["th"], 1.5e-4)
    assert rel_l2(interior(g, f["evisc"].cpu().numpy()), interior(g, n2)) <= TOL[dtype]


def o4_case(gd, g, dtype):
    from microhh_b200.synthetic import make_case
    case = make_case(gd, seed=5, noise=0.02)
    ks, ke = g.kstart, g.kend
    case["w"][:ks+1] = 0; case["w"][ke:] = 0
    case["th"] = (0.05*(case["th"] - dtype(300.))).astype(dtype)            # scalar 0 is the buoyancy
    for n in ("u", "v"):
        for sfx in ("_bot", "_top", "_gradbot", "_gradtop"):
            case[n + sfx] = np.zeros(gd.shape2d, dtype)
    case["th_gradbot"] = np.full(gd.shape2d, -0.3, dtype); case["th_gradtop"] = np.full(gd.shape2d, 0.2, dtype)
    return case


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("swadvec", ["4", "4m"])
@pytest.mark.parametrize("tbi", [0, 1, 4])
def test_full_rk3_step_order4_buoy(dtype, swadvec, tbi):
    """One full RK3 step of the 4th-order DNS configuration with swthermo = buoy registered into the fused sub-step
    (mhh_dycore_set_thermo_buoy): thermo.exec between the ghost cells and the advection (src/model.cxx:388)."""
    from microhh_b200 import dycore as D
    tb = VARIANTS[tbi]
    g, gd = grids(4, (32, 24, 16), dtype)
    case = o4_case(gd, g, dtype)
    ctx = D.Context(gd, 0)
    ones = np.ones(gd.kcells, dtype)
    ctx.set_basestate(ones, ones, 300*ones, 300*ones)
    visc = 1e-3
    f = D.Fields(ctx, case, visc=visc, svisc=visc)
    prm = D.make_params(swadvec=swadvec, swdiff="4", swthermo="buoy", surface_model=False, mbcbot=0, mbctop=0)
    oprm = ostep.default_params(); oprm.update(swadvec=swadvec, swdiff="4", swthermo="buoy", thermo_buoy=tb, visc=visc, svisc=visc, mbcbot=0, mbctop=0)
    dt = 0.01
    dy = D.Dycore(ctx, prm)
    with pytest.raises(RuntimeError, match="mhh_dycore_set_thermo_buoy"):
        dy.step(f, dt)                                                    # loud: buoy not registered (nothing has run)
    T = D.Thermo_buoy(ctx, **tb); T.register()
    c_no = copy.deepcopy(case)
    for _ in range(3):                                                    # eager, capture + replay, replay (graph on small grids)
        dy.step(f, dt)
        ostep.dycore_step(g, O.NumpyKernels(g), case, oprm, dt)
    ctx.sync()
    for n in ("u", "v", "w", "th"):
        assert rel_l2(interior(g, f[n].cpu().numpy()), interior(g, case[n])) <= 50*TOL[dtype], n
    noprm = dict(oprm); noprm.update(swthermo=None)
    for _ in range(3):
        ostep.dycore_step(g, O.NumpyKernels(g), c_no, noprm, dt)
    assert rel_l2(interior(g, c_no["w"]), interior(g, case["w"])) > 1e-5                  # the buoyancy does act (1.8e-4 here)
    T.unregister()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("swadvec", ["2i5", "2"])
def test_full_rk3_step_order2_buoy(dtype, swadvec):
    """swthermo = buoy on a 2nd-order grid with diff_2; the LES closures reject it loudly."""
    from microhh_b200 import dycore as D
    tb = VARIANTS[4]
    g, gd, case = make_pair(48, 20, 16, dtype, stretched=True)
    case["th"] = (0.05*(case["th"] - dtype(300.))).astype(dtype)
    ctx = D.Context(gd, 0)
    ctx.set_basestate(case["rhoref"], case["rhorefh"], case["thref"], case["threfh"])
    visc = 1e-2
    f = D.Fields(ctx, case, scalars=case["scalars"], visc=visc, svisc=visc)
    T = D.Thermo_buoy(ctx, **tb); T.register()
    with pytest.raises(RuntimeError, match="swthermo = buoy"):
        D.Dycore(ctx, D.make_params(swadvec=swadvec, swdiff="smag2", swthermo="buoy", surface_model=False)).step(f, 1.0)
    prm = D.make_params(swadvec=swadvec, swdiff="2", swthermo="buoy", surface_model=False)
    oprm = ostep.default_params(); oprm.update(swadvec=swadvec, swdiff="2", swthermo="buoy", thermo_buoy=tb, surface_model=False, visc=visc, svisc=visc)
    dt = 1.0
    D.Dycore(ctx, prm).step(f, dt)
    ostep.dycore_step(g, O.NumpyKernels(g), case, oprm, dt)
    ctx.sync()
    for n in ("u", "v", "w", "th"):
        assert rel_l2(interior(g, f[n].cpu().numpy()), interior(g, case[n])) <= 20*TOL[dtype], n
    T.unregister()
