"""
Damping layer and large-scale forcings on the device: `mhh_buffer_exec` / `mhh_force_exec` against the
reference's compiled kernels, and a full RK3 step with both registered into the fused sub-step (buffer.exec and force.exec run
between diff.exec and pres.exec, src/model.cxx:416-430) against the oracle stepping in the same order.
"""
import copy
import numpy as np
import pytest

from util import TOL, rel_l2, make_pair, interior, prepare_halos
from oracle import oracle as O
from oracle import step as ostep
from oracle import refbind

pytestmark = pytest.mark.gpu


def kernels(g):
    return refbind.RefKernels(g, fast=False) if refbind.available(False) else O.NumpyKernels(g)


def profiles(gd, dtype, seed=4):
    rng = np.random.default_rng(seed)
    p = lambda s=1.: (s*rng.standard_normal(gd.kcells)).astype(dtype)
    return dict(bu=p(), bv=p(), bw=p(0.1), bth=(300. + p()).astype(dtype), ug=p(), vg=p(), sls=p(1e-3), wls=p(0.01))


def apply_oracle(g, c, pr, zstart, sigma, beta, swlspres, fc, uflux, utrans, vtrans, sub_dt):
    """Buffer::exec then Force::exec on the tendencies of the case dict (src/buffer.cxx:170-205, src/force.cxx:608-700)."""
    ks, ksh = O.buffer_kstart(g, zstart)
    O.calc_buffer(g, c["ut"], c["u"], pr["bu"], g.z, zstart, beta, sigma, ks)
    O.calc_buffer(g, c["vt"], c["v"], pr["bv"], g.z, zstart, beta, sigma, ks)
    O.calc_buffer(g, c["wt"], c["w"], pr["bw"], g.zh, zstart, beta, sigma, ksh)
    O.calc_buffer(g, c["tht"], c["th"], pr["bth"], g.z, zstart, beta, sigma, ks)
    if swlspres == "uflux":
        O.force_fixed_flux(g, c["ut"], c["u"], uflux, utrans, sub_dt)
    elif swlspres == "geo":
        O.force_coriolis_2nd(g, c["ut"], c["vt"], c["u"], c["v"], pr["ug"], pr["vg"], fc, utrans, vtrans)
    O.force_ls_source(g, c["tht"], pr["sls"])
    O.force_wls_local(g, c["tht"], c["th"], pr["wls"])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("swlspres", ["geo", "uflux"])
def test_buffer_and_force_exec(dtype, swlspres):
    g, gd, case = make_pair(96, 40, 24, dtype, stretched=True)
    prepare_halos(g, case)
    rng = np.random.default_rng(8)
    for n in ("ut", "vt", "wt", "tht"):
        case[n] = (0.01*rng.standard_normal(gd.shape)).astype(dtype)
    from microhh_b200 import dycore as D
    ctx = D.Context(gd, 0)
    ctx.set_basestate(case["rhoref"], case["rhorefh"], case["thref"], case["threfh"])
    f = D.Fields(ctx, case)
    pr = profiles(gd, dtype)
    zstart = float(0.6*g.zsize)
    F = D.Forcing(ctx, f, swbuffer=True, zstart=zstart, sigma=2.5, beta=2., bufferprofs=dict(u=pr["bu"], v=pr["bv"], w=pr["bw"], th=pr["bth"]),
                  swlspres=swlspres, uflux=0.11, fc=1e-4, ug=pr["ug"], vg=pr["vg"], utrans=0.3, vtrans=-0.2,
                  ls=dict(th=pr["sls"]), wls=pr["wls"])
    F.exec_buffer(f)
    F.exec_force(f, 0.7)
    ctx.sync()
    apply_oracle(g, case, pr, zstart, 2.5, 2., swlspres, 1e-4, 0.11, 0.3, -0.2, 0.7)
    for n in ("ut", "vt", "wt", "tht"):
        assert rel_l2(f[n].cpu().numpy(), case[n]) <= 10*TOL[dtype], n


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_step_with_registered_forcing(dtype):
    """One RK3 step with the damping layer, Coriolis / geostrophic forcing, a large-scale source and subsidence inside the
    fused sub-step."""
    g, gd, case = make_pair(64, 32, 24, dtype, stretched=True)
    from microhh_b200 import dycore as D
    ctx = D.Context(gd, 0)
    ctx.set_basestate(case["rhoref"], case["rhorefh"], case["thref"], case["threfh"])
    f = D.Fields(ctx, case)
    pr = profiles(gd, dtype)
    zstart = float(0.7*g.zsize)
    F = D.Forcing(ctx, f, swbuffer=True, zstart=zstart, sigma=2., beta=2., bufferprofs=dict(u=pr["bu"], v=pr["bv"], w=pr["bw"], th=pr["bth"]),
                  swlspres="geo", fc=1e-4, ug=pr["ug"], vg=pr["vg"], ls=dict(th=pr["sls"]), wls=pr["wls"])
    F.register()
    prm = D.make_params()
    dt = 2.0
    D.Dycore(ctx, prm).step(f, dt)
    ctx.sync()
    K = kernels(g)
    oprm = ostep.default_params()
    extra = lambda c, sub_dt: apply_oracle(g, c, pr, zstart, 2., 2., "geo", 1e-4, 0., 0., 0., sub_dt)
    ostep.dycore_step(g, K, case, oprm, dt, forcing=extra)
    for n in ("u", "v", "w", "th"):
        assert rel_l2(interior(g, f[n].cpu().numpy()), interior(g, case[n])) <= 5*TOL[dtype], n
    F.unregister()
