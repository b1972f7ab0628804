"""
Advec_2i4 (reference src/advec_2i4.cxx) and Advec_2i62 (src/advec_2i62.cxx) on the device: tendencies and CFL number through
mhh_advec_exec / mhh_advec_get_cfl with swadvec = 24 / 262, and full RK3 steps inside the fused sub-step (cases/gabls4s3 is
2i4 + smag2 + dry; cases/weisman_klemp advects with 2i62), against the oracle (pinned bit for bit to the compiled reference in
tests/test_oracle_vs_ref.py).  Tolerances: relative L2 <= 1e-12 (fp64), <= 1e-5 (fp32).
"""
import numpy as np
import pytest

from util import TOL, rel_l2, interior, stretched_z
from oracle import oracle as O
from oracle import step as ostep

pytestmark = pytest.mark.gpu

DTYPES = [np.float64, np.float32]
GC = {"2i4": (2, 2, 2), "2i62": (3, 3, 1)}


def pair(shape, dtype, gc, ns=1, anel=True, seed=4):
    from microhh_b200.grid import GridData
    from microhh_b200.synthetic import make_case
    z = stretched_z(shape[2], 3200.)
    g = O.Grid(*shape, 3200., 3200., 3200., *gc, dtype, z=z)
    gd = GridData(*shape, 3200., 3200., 3200., *gc, dtype, z=z)
    return g, gd, make_case(gd, seed=seed, anelastic=anel, ns=ns)


def setup(gd, case, **fkw):
    from microhh_b200 import dycore as D
    ctx = D.Context(gd, 0)
    ctx.set_basestate(case["rhoref"], case["rhorefh"], case["thref"], case["threfh"])
    return D, ctx, D.Fields(ctx, case, scalars=case["scalars"], **fkw)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("scheme", ["2i4", "2i62"])
@pytest.mark.parametrize("shape", [(32, 16, 12), (72, 10, 9), (24, 1, 8)])
@pytest.mark.parametrize("anel", [False, True])
def test_advec_2i4_2i62(dtype, scheme, shape, anel):
    g, gd, case = pair(shape, dtype, GC[scheme], anel=anel)
    rng = np.random.default_rng(9)
    for n in ("u", "v", "w", "th", "ut", "vt", "wt", "tht"):
        case[n] = (case[n] + 0.1*rng.standard_normal(gd.shape)).astype(dtype)       # ghost cells and tendencies too
    D, ctx, f = setup(gd, case)
    A = D.Advec(ctx, scheme)
    A.exec(f)
    K = O.NumpyKernels(g)
    ref = {n: case[n].copy() for n in ("ut", "vt", "wt", "tht")}
    a = (case["u"], case["v"], case["w"], case["rhoref"], case["rhorefh"])
    getattr(K, f"advec_{scheme}_u")(ref["ut"], *a); getattr(K, f"advec_{scheme}_v")(ref["vt"], *a); getattr(K, f"advec_{scheme}_w")(ref["wt"], *a)
    getattr(K, f"advec_{scheme}_s")(ref["tht"], case["th"], *a)
    for n in ref:
        got = f[n].cpu().numpy()
        assert rel_l2(got, ref[n]) <= TOL[dtype], n
        assert np.array_equal(got[:g.kstart], case[n][:g.kstart]) and np.array_equal(got[g.kend:], case[n][g.kend:]), n   # ghost levels untouched
    assert np.array_equal(f["wt"].cpu().numpy()[g.kstart], case["wt"][g.kstart])
    assert not np.array_equal(f["ut"].cpu().numpy(), case["ut"])
    cfl = A.get_cfl(f, 3.0)
    assert abs(cfl - getattr(K, f"advec_{scheme}_cfl")(case["u"], case["v"], case["w"], 3.0)) <= 10*TOL[dtype]*cfl
    # the ghost cells each scheme asks for are checked loudly
    if scheme == "2i4":
        g1, gd1, c1 = pair(shape, dtype, (2, 2, 1))
        D1, ctx1, f1 = setup(gd1, c1)
        with pytest.raises(RuntimeError, match="kgc >= 2"):
            D1.Advec(ctx1, "2i4").exec(f1)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("scheme,gc,swdiff,thermo", [("2i4", (2, 2, 2), "smag2", "dry"), ("2i4", (2, 2, 2), "2", "dry"), ("2i4", (4, 2, 2), "smag2", None),
                                                     ("2i62", (3, 3, 1), "smag2", "dry"), ("2i62", (4, 3, 2), "smag2", "dry"), ("2i62", (3, 3, 2), "2", None)])
def test_full_rk3_step_2i4_2i62(dtype, scheme, gc, swdiff, thermo):
    """One full RK3 step inside the fused sub-step: the advection alone by adv2i_*_kernel, then the diffusion (+ buoyancy) kernels;
    two scalars, the second one flux-limited where the grid has kgc = 2 and the scheme is 2i62."""
    shape = (64, 24, 16)
    g, gd, case = pair(shape, dtype, gc, ns=2)
    smag = swdiff == "smag2"
    visc = 1e-5 if smag else 1e-2
    lim = ("s1",) if gc[2] == 2 and scheme == "2i62" else ()
    oprm = ostep.default_params(); oprm.update(swadvec=scheme, swdiff=swdiff, swthermo=thermo, surface_model=smag, visc=visc, svisc=visc, fluxlimit_list=lim)
    dt = 2.0
    from microhh_b200 import dycore as D
    ctx = D.Context(gd, 0)
    ctx.set_basestate(case["rhoref"], case["rhorefh"], case["thref"], case["threfh"])
    f = D.Fields(ctx, case, scalars=case["scalars"], visc=visc, svisc=visc, fluxlimit_list=lim)
    prm = D.make_params(swadvec=scheme, swdiff=swdiff, swthermo=thermo or "0", surface_model=smag, ns=2)
    ctx.profile_start()
    D.Dycore(ctx, prm).step(f, dt)
    prof = ctx.profile_stop()
    ctx.sync()
    assert "adv2i_uvw_kernel" in prof and ("advec_s_lim_kernel" in prof) == bool(lim), sorted(prof)
    ostep.dycore_step(g, O.NumpyKernels(g), case, oprm, dt)
    names = ("u", "v", "w", "th", "s1")
    tol = 20*TOL[dtype] if dtype == np.float64 else 5*TOL[dtype]
    for n in names:
        assert rel_l2(interior(g, f[n].cpu().numpy()), interior(g, case[n])) <= tol, n
