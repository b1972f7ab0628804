"""
Thermo_moist on the device (reference src/thermo_moist.cxx, include/thermo_moist_functions.h): the hydrostatic base state
integrated on the GPU, the mean profiles, the buoyancy tendency through the saturation adjustment, the b / ql / N2 diagnostics,
the surface buoyancy helpers and the fused sub-steps with swthermo = moist, through the C ABI against the oracle (pinned bit
for bit to the compiled reference in tests/test_oracle_vs_ref.py).

Tolerances: relative L2 <= 1e-12 (fp64) / <= 1e-5 (fp32) on tendencies as they occur in a step and on the prognostic fields.
Two places get a stated, looser bound because the formula itself amplifies rounding: the buoyancy g (thv - thvref) / thvref
subtracts two numbers of ~300 K that differ by ~1 K (amplification thvref / |thv - thvref| ~ 1e2..1e3), and the base state is
a kmax-long chain of exp / pow calls whose device versions differ from glibc's by an ulp per call.
"""
import copy
import numpy as np
import pytest

from util import TOL, rel_l2, interior, make_moist_pair
from oracle import oracle as O
from oracle import step as ostep

pytestmark = pytest.mark.gpu

DTYPES = [np.float64, np.float32]
RANGES = dict(pref=(-1, 1), prefh=(0, 1), rhoref=(0, 0), thvref=(0, 0), exnref=(0, 0), rhorefh=(0, 1), thvrefh=(0, 1), exnrefh=(0, 1))


def setup(gd, case, **fkw):
    from microhh_b200 import dycore as D
    ctx = D.Context(gd, 0)
    ctx.set_basestate(case["rhoref"], case["rhorefh"], case["thref"], case["threfh"])
    f = D.Fields(ctx, case, scalars=case["scalars"], **fkw)
    return D, ctx, f


def profiles_close(g, got, ref, tol):
    for n, (lo, hi) in RANGES.items():
        sl = slice(g.kstart + lo, g.kend + hi)
        assert rel_l2(got[n][sl], ref[n][sl]) <= tol, (n, got[n][sl], ref[n][sl])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cold", [False, True])
def test_moist_base_state(dtype, cold):
    """calc_base_state on the device (one lane, serial in k) == the oracle; set / get of the profiles is exact."""
    g, gd, case, pbot = make_moist_pair(32, 16, 24, dtype, cold=cold)
    D, ctx, f = setup(gd, case)
    T = D.Thermo_moist(ctx, f, pbot)
    with pytest.raises(RuntimeError, match="no base state"):
        T.exec(f)
    thl0, qt0 = case["moist_ref"]
    T.calc_base_state(thl0, qt0)
    profiles_close(g, T.get_profiles(), case["moist_bs"], 1e-13 if dtype == np.float64 else 2e-6)
    assert T.nonconverged() == 0
    with pytest.raises(RuntimeError, match="both or neither"):
        T.set_profiles(prefh=case["moist_bs"]["prefh"])                     # a pressure profile travels with its exner function
    T.set_profiles(**case["moist_bs"])
    got = T.get_profiles()
    for n in case["moist_bs"]:
        assert np.array_equal(got[n], case["moist_bs"][n]), n


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cold", [False, True])
@pytest.mark.parametrize("shape", [(32, 16, 24), (72, 10, 12)])
def test_moist_exec_and_fields(dtype, cold, shape):
    """Thermo_moist::exec without base-state update and the get_thermo_field / surface helpers, on the oracle's profiles."""
    import torch
    g, gd, case, pbot = make_moist_pair(*shape, dtype, cold=cold)
    bs = case["moist_bs"]
    K = O.NumpyKernels(g)
    rng = np.random.default_rng(4)
    seed = rng.standard_normal(gd.shape).astype(dtype)
    g64, gd64, c64, _ = make_moist_pair(*shape, np.float64, cold=cold)
    K64 = O.NumpyKernels(g64)
    b64 = {n: a.astype(np.float64) for n, a in bs.items()}
    res = {}
    for name, wt0 in (("seeded", seed), ("zero", np.zeros(gd.shape, dtype))):
        case["wt"] = wt0.copy()
        D, ctx, f = setup(gd, case)
        T = D.Thermo_moist(ctx, f, pbot, swupdatebasestate=False)
        T.set_profiles(**bs)
        T.exec(f)
        ref = wt0.copy()
        K.thermo_moist_buoyancy_tend_2nd(ref, case["thl"], case["qt"], bs["prefh"], bs["thvrefh"])
        got = f["wt"].cpu().numpy()
        assert np.array_equal(got[:g.kstart+1], wt0[:g.kstart+1]) and np.array_equal(got[g.kend:], wt0[g.kend:])   # wall level + ghosts untouched
        if dtype == np.float64:
            assert rel_l2(got, ref) <= (TOL[dtype] if name == "seeded" else 100*TOL[dtype]), name
        else:
            # fp32: as close to the fp64 evaluation of the same inputs as the reference's own fp32 kernels are
            t = wt0.astype(np.float64)
            K64.thermo_moist_buoyancy_tend_2nd(t, case["thl"].astype(np.float64), case["qt"].astype(np.float64), b64["prefh"], b64["thvrefh"])
            e_gpu, e_ref = rel_l2(got, t), rel_l2(ref, t)
            assert e_gpu <= 1.25*e_ref + 1e-6, (name, e_gpu, e_ref)
            if name == "seeded":
                assert rel_l2(got, ref) <= TOL[dtype]
        res[name] = got
        assert T.nonconverged() == 0
    assert float(np.abs(interior(g, res["zero"], g.kstart+1)).max()) > 1e-3                    # there is a buoyancy
    # diagnostics: ql is a small difference of two ~1e-2 numbers, N2 and b are plain
    out = {}
    for name in ("b", "ql", "N2"):
        f["evisc"].zero_()
        T.get_thermo_field(f["evisc"], name, f)
        out[name] = f["evisc"].cpu().numpy().copy()
    ref = {n: np.zeros(gd.shape, dtype) for n in out}
    thv_full = bs["thvref"] + (bs["thvref"] == 0)*dtype(300.)
    T.set_profiles(thvref=thv_full)                                                             # calc_buoyancy divides by thvref at the ghost levels too
    f["evisc"].zero_(); T.get_thermo_field(f["evisc"], "b", f); out["b"] = f["evisc"].cpu().numpy().copy()
    K.thermo_moist_buoyancy(ref["b"], case["thl"], case["qt"], bs["pref"], thv_full)
    K.thermo_moist_liquid_water(ref["ql"], case["thl"], case["qt"], bs["pref"])
    K.thermo_moist_N2(ref["N2"], case["thl"], bs["thvref"])
    cols = (slice(None), slice(g.jstart, g.jend), slice(g.istart, g.iend))
    assert rel_l2(out["b"][cols], ref["b"][cols]) <= (100*TOL[dtype] if dtype == np.float64 else 5e-4)
    assert rel_l2(interior(g, out["ql"]), interior(g, ref["ql"])) <= (100*TOL[dtype] if dtype == np.float64 else 1e-3)
    assert rel_l2(interior(g, out["N2"]), interior(g, ref["N2"])) <= TOL[dtype]
    assert (interior(g, out["ql"]) > 0).mean() > 0.05
    # surface buoyancy
    thlbot = (case["thl"][g.kstart] + dtype(0.5)).astype(dtype); qtbot = (case["qt"][g.kstart]*dtype(1.1)).astype(dtype)
    f["thl_bot"].copy_(torch.from_numpy(thlbot)); f["qt_bot"].copy_(torch.from_numpy(qtbot))
    f["p"].zero_()
    T.get_buoyancy_surf(f["p"], f["u_bot"], f)
    rb = np.zeros(gd.shape, dtype); rbot = np.zeros(gd.shape2d, dtype)
    K.thermo_moist_buoyancy_bot(rb, rbot, case["thl"], thlbot, case["qt"], qtbot, thv_full, bs["thvrefh"])
    stol = 100*TOL[dtype] if dtype == np.float64 else 5e-4
    assert rel_l2(f["p"].cpu().numpy(), rb) <= stol and rel_l2(f["u_bot"].cpu().numpy(), rbot) <= stol
    T.get_buoyancy_fluxbot(f["v_bot"], f)
    rf = np.zeros(gd.shape2d, dtype)
    K.thermo_moist_buoyancy_fluxbot(rf, case["thl"], case["thl_fluxbot"], case["qt"], case["qt_fluxbot"], bs["thvrefh"])
    assert rel_l2(f["v_bot"].cpu().numpy(), rf) <= 10*TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
def test_moist_exec_update_basestate(dtype):
    """swupdatebasestate: mean profiles + base state on the device inside exec (no host round trip)."""
    g, gd, case, pbot = make_moist_pair(48, 20, 24, dtype)
    # shift the state so that the updated base state differs from the initial one
    case["thl"] = (case["thl"] + dtype(0.7)).astype(dtype)
    K = O.NumpyKernels(g)
    D, ctx, f = setup(gd, case)
    T = D.Thermo_moist(ctx, f, pbot, swupdatebasestate=True)
    T.set_profiles(**case["moist_bs"])
    T.exec(f)
    c = dict(thl=case["thl"], qt=case["qt"], wt=np.zeros(gd.shape, dtype), moist_bs=case["moist_bs"])
    ostep.thermo_moist_exec(K, c, dict(pbot=pbot, swupdatebasestate=True))
    assert not np.array_equal(c["moist_bs"]["thvrefh"], case["moist_bs"]["thvrefh"])
    profiles_close(g, T.get_profiles(), c["moist_bs"], 1e-13 if dtype == np.float64 else 2e-6)
    # the device's thvrefh differs from the oracle's by the rounding of the exp / pow chain (~1e-15 relative); the buoyancy
    # amplifies that by thvref / |thv - thvref|
    assert rel_l2(f["wt"].cpu().numpy(), c["wt"]) <= (1e-10 if dtype == np.float64 else 2e-3)
    assert T.nonconverged() == 0


def truth64(case32, shape, oprm, dt, pbot, **kw):
    g64, gd64, c64, _ = make_moist_pair(*shape, np.float64, **kw)
    for k, a in case32.items():
        if isinstance(a, np.ndarray):
            c64[k] = a.astype(np.float64)
    thl0, qt0 = (a.astype(np.float64) for a in case32["moist_ref"])
    c64["moist_bs"] = O.moist_base_state(g64, thl0, qt0, pbot)
    ostep.dycore_step(g64, O.NumpyKernels(g64), c64, copy.deepcopy(oprm), dt)
    return g64, c64


@pytest.mark.parametrize("dtype,igc", [(np.float64, 3), (np.float64, 4), (np.float32, 4)])
@pytest.mark.parametrize("swadvec,swdiff,update", [("2i5", "smag2", True), ("2", "smag2", True), ("2i5", "smag2", False), ("2i5", "2", True)])
def test_full_rk3_step_moist(dtype, igc, swadvec, swdiff, update):
    """One full RK3 step with swthermo = moist registered into the fused sub-step (bomex-like state, anelastic base state):
    closure N2 from thvref, base state updated in every sub-step, buoyancy through the saturation adjustment."""
    from microhh_b200 import dycore as D
    shape = (96, 24, 32)
    g, gd, case, pbot = make_moist_pair(*shape, dtype, igc=igc)
    smag = swdiff == "smag2"
    visc = 1e-5 if smag else 1e-2
    oprm = ostep.default_params(); oprm.update(swadvec=swadvec, swdiff=swdiff, swthermo="moist", surface_model=smag, visc=visc, svisc=visc,
                                               thermo_moist=dict(pbot=pbot, swupdatebasestate=update))
    dt = 2.0
    names = ["u", "v", "w", "thl", "qt"]
    if dtype == np.float32:
        g64, c64 = truth64(case, shape, oprm, dt, pbot, igc=igc)
    Dm, ctx, f = setup(gd, case, visc=visc, svisc=visc)
    prm = D.make_params(swadvec=swadvec, swdiff=swdiff, swthermo="moist", surface_model=smag, ns=2)
    dy = D.Dycore(ctx, prm)
    with pytest.raises(RuntimeError, match="mhh_dycore_set_thermo_moist"):
        dy.step(f, dt)
    T = D.Thermo_moist(ctx, f, pbot, swupdatebasestate=update)
    T.calc_base_state(*case["moist_ref"])
    T.register()
    dy.step(f, dt)
    ostep.dycore_step(g, O.NumpyKernels(g), case, oprm, dt)
    ctx.sync()
    assert T.nonconverged() == 0
    if dtype == np.float32:
        for n in names:
            t = interior(g64, c64[n])
            e_gpu = rel_l2(interior(g, f[n].cpu().numpy()), t); e_ref = rel_l2(interior(g, case[n]), t)
            assert e_gpu <= 1.25*e_ref + 1e-6, (n, e_gpu, e_ref)
        tol = 5*TOL[dtype]
    else:
        tol = 20*TOL[dtype]
    for n in names:
        assert rel_l2(interior(g, f[n].cpu().numpy()), interior(g, case[n])) <= tol, n
    if update:
        profiles_close(g, T.get_profiles(), case["moist_bs"], 1e-12 if dtype == np.float64 else 2e-6)
    T.unregister()
