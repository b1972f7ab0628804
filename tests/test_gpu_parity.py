"""
Parity tests proper: the CUDA path, called through the C ABI (libmhhb200.so), against the
oracle on the same seeded inputs.  Tolerances are BASELINE.json's: relative L2 <= 1e-12 in
fp64 and <= 1e-5 in fp32; copies (boundary_cyclic) are bit-exact.
"""
import copy
import numpy as np
import pytest

from util import TOL, rel_l2, make_pair, interior, prepare_halos
from oracle import oracle as O
from oracle import step as ostep

pytestmark = pytest.mark.gpu

DTYPES = [np.float64, np.float32]
SHAPES = [(32, 16, 12), (48, 40, 24), (20, 12, 8)]


def gpu_setup(gd, case, ns=1):
    import torch
    from microhh_b200 import dycore as D
    ctx = D.Context(gd, 0)
    ctx.set_basestate(case["rhoref"], case["rhorefh"], case["thref"], case["threfh"])
    f = D.Fields(ctx, case, scalars=case["scalars"])
    prm = D.make_params(ns=ns)
    return D, ctx, f, prm


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(16, 12, 8), (32, 1, 8)])
@pytest.mark.parametrize("edge", [0, 1, 2])
def test_boundary_cyclic_bitexact(dtype, shape, edge):
    g, gd, case = make_pair(*shape, dtype)
    rng = np.random.default_rng(5)
    a = rng.standard_normal(gd.shape).astype(dtype)
    case["u"] = a.copy()
    D, ctx, f, prm = gpu_setup(gd, case)
    D.Boundary_cyclic(ctx).exec(f["u"], edge)
    O.boundary_cyclic(g, a, edge)
    assert np.array_equal(f["u"].cpu().numpy(), a)
    a2 = rng.standard_normal(gd.shape2d).astype(dtype)
    f["z0m"].copy_(__import__("torch").from_numpy(a2))
    D.Boundary_cyclic(ctx).exec_2d(f["z0m"])
    O.boundary_cyclic_2d(g, a2)
    assert np.array_equal(f["z0m"].cpu().numpy(), a2)


@pytest.mark.parametrize("dtype", DTYPES)
def test_ghost_cells(dtype):
    g, gd, case = make_pair(16, 12, 8, dtype, stretched=True)
    D, ctx, f, prm = gpu_setup(gd, case)
    B = D.Boundary(ctx)
    for bc in (0, 1):
        a = case["u"].copy()
        f["u"].copy_(__import__("torch").from_numpy(a))
        B.set_ghost_cells_field(f["u"], bc, f["u_gradbot"], f["u_gradbot"], bc, f["u_gradtop"], f["u_gradtop"])
        O.ghost_cells_bot_2nd(g, a, bc, case["u_gradbot"], case["u_gradbot"])
        O.ghost_cells_top_2nd(g, a, bc, case["u_gradtop"], case["u_gradtop"])
        assert rel_l2(f["u"].cpu().numpy(), a) <= TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("anel", [False, True])
def test_advec_2i5(dtype, shape, anel):
    g, gd, case = make_pair(*shape, dtype, stretched=True, anelastic=anel)
    prepare_halos(g, case)
    D, ctx, f, prm = gpu_setup(gd, case)
    D.Advec(ctx, "2i5").exec(f)
    rr, rh = case["rhoref"], case["rhorefh"]
    ref = {n: g.field() for n in ("ut", "vt", "wt", "tht")}
    O.advec_2i5_u(g, ref["ut"], case["u"], case["v"], case["w"], rr, rh)
    O.advec_2i5_v(g, ref["vt"], case["u"], case["v"], case["w"], rr, rh)
    O.advec_2i5_w(g, ref["wt"], case["u"], case["v"], case["w"], rr, rh)
    O.advec_2i5_s(g, ref["tht"], case["th"], case["u"], case["v"], case["w"], rr, rh)
    for n in ref:
        assert rel_l2(f[n].cpu().numpy(), ref[n]) <= TOL[dtype], n
    cfl = D.Advec(ctx, "2i5").get_cfl(f, 3.0)
    assert abs(cfl - float(O.advec_2i5_cfl(g, case["u"], case["v"], case["w"], 3.0))) <= 10*TOL[dtype]*cfl


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("surface", [True, False])
def test_diff_smag2(dtype, shape, surface):
    g, gd, case = make_pair(*shape, dtype, stretched=True, anelastic=True)
    prepare_halos(g, case)
    D, ctx, f, _ = gpu_setup(gd, case)
    prm = D.make_params(surface_model=surface)
    diff = D.Diff(ctx, prm)
    # exec_viscosity, N2 derived from th
    diff.exec_viscosity(f)
    ev = g.field(); N2 = g.field()
    O.diff_strain2(g, ev, case["u"], case["v"], case["w"], case["dudz_mo"], case["dvdz_mo"], surface)
    O.thermo_dry_N2(g, N2, case["th"], case["thref"])
    O.diff_evisc(g, ev, N2, case["dbdz_mo"], case["z0m"], 0.23, 1./3., surface)
    k0 = g.kstart - (0 if surface else 1); k1 = g.kend + (0 if surface else 1)
    got = f["evisc"].cpu().numpy()
    assert rel_l2(got[k0:k1], ev[k0:k1]) <= TOL[dtype]
    # exec_viscosity with an externally supplied N2 field
    import torch
    n2_dev = torch.zeros_like(f["evisc"])
    D.Thermo_dry(ctx).get_thermo_field_N2(n2_dev, f)
    assert rel_l2(interior(g, n2_dev.cpu().numpy()), interior(g, N2)) <= TOL[dtype]
    f["evisc"].zero_()
    diff.exec_viscosity(f, n2_dev)
    assert rel_l2(f["evisc"].cpu().numpy()[k0:k1], ev[k0:k1]) <= TOL[dtype]
    # exec
    f["evisc"].copy_(torch.from_numpy(ev))
    diff.exec(f)
    rr, rh = case["rhoref"], case["rhorefh"]
    ref = {n: g.field() for n in ("ut", "vt", "wt", "tht")}
    O.diff_u(g, ref["ut"], case["u"], case["v"], case["w"], ev, case["u_fluxbot"], case["u_fluxtop"], rr, rh, 1e-5, surface)
    O.diff_v(g, ref["vt"], case["u"], case["v"], case["w"], ev, case["v_fluxbot"], case["v_fluxtop"], rr, rh, 1e-5, surface)
    O.diff_w(g, ref["wt"], case["u"], case["v"], case["w"], ev, rr, rh, 1e-5)
    O.diff_c(g, ref["tht"], case["th"], ev, case["th_fluxbot"], case["th_fluxtop"], rr, rh, 1./3., 1e-5, surface)
    for n in ref:
        assert rel_l2(f[n].cpu().numpy(), ref[n]) <= TOL[dtype], n
    dn = diff.get_dn(f, 2.0)
    assert abs(dn - 2.0*float(O.diff_dnmul(g, ev, 1./3.))) <= 10*TOL[dtype]*dn


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("anel", [False, True])
def test_advec_2_and_diff_2(dtype, shape, anel):
    """Advec_2 (src/advec_2.cxx) and Diff_2 (src/diff_2.cxx) kernels, cfl and dn against the oracle."""
    g, gd, case = make_pair(*shape, dtype, stretched=True, anelastic=anel)
    prepare_halos(g, case)
    D, ctx, f, prm = gpu_setup(gd, case)
    D.Advec(ctx, "2").exec(f)
    rr, rh = case["rhoref"], case["rhorefh"]
    ref = {n: g.field() for n in ("ut", "vt", "wt", "tht")}
    O.advec_2_u(g, ref["ut"], case["u"], case["v"], case["w"], rr, rh)
    O.advec_2_v(g, ref["vt"], case["u"], case["v"], case["w"], rr, rh)
    O.advec_2_w(g, ref["wt"], case["u"], case["v"], case["w"], rr, rh)
    O.advec_2_s(g, ref["tht"], case["th"], case["u"], case["v"], case["w"], rr, rh)
    for n in ref:
        assert rel_l2(f[n].cpu().numpy(), ref[n]) <= TOL[dtype], n
    cfl = D.Advec(ctx, "2").get_cfl(f, 3.0)
    assert abs(cfl - float(O.advec_2_cfl(g, case["u"], case["v"], case["w"], 3.0))) <= 10*TOL[dtype]*cfl
    # Diff_2 on top of the advection tendencies (the kernels accumulate)
    f.visc, f.svisc = 0.5, 0.7
    f._build()
    D.Diff_2(ctx).exec(f)
    O.diff_2_c(g, ref["ut"], case["u"], 0.5); O.diff_2_c(g, ref["vt"], case["v"], 0.5)
    O.diff_2_w(g, ref["wt"], case["w"], 0.5); O.diff_2_c(g, ref["tht"], case["th"], 0.7)
    for n in ref:
        assert rel_l2(f[n].cpu().numpy(), ref[n]) <= TOL[dtype], n
    dn = D.Diff_2(ctx).get_dn(f, 2.0)
    assert abs(dn - 2.0*O.diff_2_dnmul(g, 0.7)) <= 10*TOL[dtype]*dn


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("swadvec,swdiff", [("2", "2"), ("2", "smag2"), ("2i5", "2")])
def test_full_rk3_step_scheme_combinations(dtype, swadvec, swdiff):
    """Full RK3 step with Advec_2 / Diff_2 in every combination with the LES schemes."""
    g, gd, case = make_pair(32, 24, 16, dtype, stretched=True, anelastic=True)
    D, ctx, f, _ = gpu_setup(gd, case)
    f.visc, f.svisc = 0.5, 0.7
    f._build()
    prm = D.make_params(swadvec=swadvec, swdiff=swdiff)
    oprm = ostep.default_params(); oprm.update(swadvec=swadvec, swdiff=swdiff, visc=0.5, svisc=0.7)
    D.Dycore(ctx, prm).step(f, 2.0)
    ostep.dycore_step(g, O.NumpyKernels(g), case, oprm, 2.0)
    ctx.sync()
    for n in ("u", "v", "w", "th"):
        assert rel_l2(interior(g, f[n].cpu().numpy()), interior(g, case[n])) <= TOL[dtype], n


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(32, 16, 12), (24, 1, 8), (20, 12, 6)])
@pytest.mark.parametrize("stretched", [False, True])
def test_advec_4_and_diff_4(dtype, shape, stretched):
    """Advec_4 (src/advec_4.cxx) and Diff_4 (src/diff_4.cxx) on a 4th-order grid, 3-D and 2-D (jtot = 1)."""
    import torch
    from util import stretched_z
    from microhh_b200 import dycore as D
    from microhh_b200.grid import GridData
    it, jt, kt = shape
    z = stretched_z(kt, 3200.) if stretched else None
    g = O.Grid(it, jt, kt, 3200., 3200., 3200., 3, 3, 3, dtype, z=z, order=4)
    gd = GridData(it, jt, kt, 3200., 3200., 3200., 3, 3, 3, dtype, z=z, order=4)
    rng = np.random.default_rng(3)
    fld = lambda: rng.standard_normal(gd.shape).astype(dtype)
    case = dict(u=fld(), v=fld(), w=fld(), th=fld(), ut=fld(), vt=fld(), wt=fld(), tht=fld())
    ctx = D.Context(gd, 0)
    ones = np.ones(gd.kcells, dtype)
    ctx.set_basestate(ones, ones, 300*ones, 300*ones)
    f = D.Fields(ctx, case, visc=0.7, svisc=1.3)
    D.Advec(ctx, "4").exec(f)
    ref = {n: case[n].copy() for n in ("ut", "vt", "wt", "tht")}
    O.advec_4_u(g, ref["ut"], case["u"], case["v"], case["w"]); O.advec_4_v(g, ref["vt"], case["u"], case["v"], case["w"])
    O.advec_4_w(g, ref["wt"], case["u"], case["v"], case["w"]); O.advec_4_s(g, ref["tht"], case["th"], case["u"], case["v"], case["w"])
    for n in ref:
        assert rel_l2(f[n].cpu().numpy(), ref[n]) <= TOL[dtype], n
    cfl = D.Advec(ctx, "4").get_cfl(f, 3.0)
    assert abs(cfl - float(O.advec_4_cfl(g, case["u"], case["v"], case["w"], 3.0))) <= 10*TOL[dtype]*cfl
    D.Diff_4(ctx).exec(f)
    O.diff_4_c(g, ref["ut"], case["u"], 0.7); O.diff_4_c(g, ref["vt"], case["v"], 0.7)
    O.diff_4_w(g, ref["wt"], case["w"], 0.7); O.diff_4_c(g, ref["tht"], case["th"], 1.3)
    for n in ref:
        assert rel_l2(f[n].cpu().numpy(), ref[n]) <= TOL[dtype], n
    # a 2nd-order grid refuses the 4th-order schemes loudly
    g2, gd2, case2 = make_pair(16, 12, 8, dtype)
    D2, ctx2, f2, _ = gpu_setup(gd2, case2)
    with pytest.raises(D.MhhError):
        D.Advec(ctx2, "4").exec(f2)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(32, 24, 16), (48, 40, 12), (24, 1, 8), (64, 32, 20)])
@pytest.mark.parametrize("stretched", [False, True])
def test_pres_4_exec(dtype, shape, stretched):
    """Pres_4 (src/pres_4.cxx): 4th-order rhs, transforms, 7-band solve, ghost levels, tendency correction == oracle;
    the corrected velocity is divergence-free under the 4th-order divergence."""
    from util import stretched_z
    from microhh_b200 import dycore as D
    from microhh_b200.grid import GridData
    it, jt, kt = shape
    z = stretched_z(kt, 2.) if stretched else None
    g = O.Grid(it, jt, kt, 6., 4., 2., 3, 3, 3, dtype, z=z, order=4)
    gd = GridData(it, jt, kt, 6., 4., 2., 3, 3, 3, dtype, z=z, order=4)
    rng = np.random.default_rng(1)
    fld = lambda: rng.standard_normal(gd.shape).astype(dtype)
    u, v, w = fld(), fld(), fld()
    if jt == 1:
        v[:] = 0
    ks, ke = g.kstart, g.kend
    w[ks] = 0; w[ke] = 0; w[ks-1] = -w[ks+1]; w[ke+1] = -w[ke-1]
    for a in (u, v, w):
        O.boundary_cyclic(g, a)
    case = dict(u=u, v=v, w=w, th=fld())
    for n in ("ut", "vt", "wt"):
        t = np.zeros(gd.shape, dtype)
        interior(g, t)[...] = 0.1*rng.standard_normal((gd.kmax, gd.jmax, gd.imax))
        case[n] = t
    case["wt"][:ks+1] = 0
    if jt == 1:
        case["vt"][:] = 0
    ctx = D.Context(gd, 0)
    ones = np.ones(gd.kcells, dtype)
    ctx.set_basestate(ones, ones, 300*ones, 300*ones)
    f = D.Fields(ctx, case)
    dt = 0.5
    D.Pres(ctx, 4).exec(f, dt)
    P = O.Pres4(g)
    ref = {n: case[n].copy() for n in ("ut", "vt", "wt")}
    p = g.field()
    P.exec(p, u, v, w, ref["ut"], ref["vt"], ref["wt"], dt)
    ptol = 200*TOL[dtype]
    got_p = f["p"].cpu().numpy()
    # interior levels incl. lateral ghosts; the two ghost levels at either wall on the interior columns (the 2-D cyclic fill
    # of the reference leaves the lateral ghosts of the ghost levels untouched, src/boundary_cyclic.cxx:427-441)
    assert rel_l2(got_p[ks:ke], p[ks:ke]) <= ptol
    assert rel_l2(interior(g, got_p, ks-2, ke+2), interior(g, p, ks-2, ke+2)) <= ptol
    for n in ref:
        if jt == 1 and n == "vt":
            continue
        assert rel_l2(interior(g, f[n].cpu().numpy()), interior(g, ref[n])) <= ptol, n
    un = {c: case[c] + dtype(dt)*f[c + "t"].cpu().numpy() for c in "uvw"}
    un["w"][ks-1] = -un["w"][ks+1]; un["w"][ke+1] = -un["w"][ke-1]
    for c in "uvw":
        O.boundary_cyclic(g, un[c])
    div0 = float(P.divergence(u, v, w))
    assert float(P.divergence(un["u"], un["v"], un["w"])) <= (1e-11 if dtype == np.float64 else 2e-3)*div0
    assert abs(D.Pres(ctx, 4).check_divergence(f) - div0) <= 100*TOL[dtype]*div0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(32, 16, 12), (24, 1, 8), (20, 12, 6)])
@pytest.mark.parametrize("stretched", [False, True])
def test_advec_4m(dtype, shape, stretched):
    """Advec_4m (src/advec_4m.cxx:51-415), the fully conservative 4th-order scheme cases/moser180 ships with: tendencies of
    u, v, w and a scalar and the CFL number through mhh_advec_exec / mhh_advec_get_cfl with swadvec = 41."""
    from util import stretched_z
    from microhh_b200 import dycore as D
    from microhh_b200.grid import GridData
    it, jt, kt = shape
    z = stretched_z(kt, 2.) if stretched else None
    g = O.Grid(it, jt, kt, 6.28, 3.14, 2., 3, 3, 3, dtype, z=z, order=4)
    gd = GridData(it, jt, kt, 6.28, 3.14, 2., 3, 3, 3, dtype, z=z, order=4)
    rng = np.random.default_rng(7)
    fld = lambda: rng.standard_normal(gd.shape).astype(dtype)
    case = dict(u=fld(), v=fld(), w=fld(), th=fld(), ut=fld(), vt=fld(), wt=fld(), tht=fld())
    ctx = D.Context(gd, 0)
    ones = np.ones(gd.kcells, dtype)
    ctx.set_basestate(ones, ones, 300*ones, 300*ones)
    f = D.Fields(ctx, case, visc=0.7, svisc=1.3)
    D.Advec(ctx, "4m").exec(f)
    ref = {n: case[n].copy() for n in ("ut", "vt", "wt", "tht")}
    O.advec_4m_u(g, ref["ut"], case["u"], case["v"], case["w"]); O.advec_4m_v(g, ref["vt"], case["u"], case["v"], case["w"])
    O.advec_4m_w(g, ref["wt"], case["u"], case["v"], case["w"]); O.advec_4m_s(g, ref["tht"], case["th"], case["u"], case["v"], case["w"])
    for n in ref:
        assert rel_l2(f[n].cpu().numpy(), ref[n]) <= TOL[dtype], n
        assert np.array_equal(f[n].cpu().numpy()[:g.kstart], case[n][:g.kstart])        # ghost levels untouched
    cfl = D.Advec(ctx, "4m").get_cfl(f, 3.0)
    assert abs(cfl - float(O.advec_4m_cfl(g, case["u"], case["v"], case["w"], 3.0))) <= 10*TOL[dtype]*cfl


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,mbc", [((32, 24, 16), 0), ((24, 1, 12), 0), ((32, 16, 12), 1)])
@pytest.mark.parametrize("swadvec", ["4", "4m"])
def test_full_rk3_step_order4(dtype, shape, mbc, swadvec):
    """The 4th-order DNS configuration (moser180 / taylorgreen style: advec_4 or advec_4m + diff_4 + pres_4, 4th-order ghost cells,
    no thermo): one full RK3 step == the oracle in the reference's call order; no-slip (Dirichlet) and free-slip walls."""
    from util import stretched_z
    from microhh_b200 import dycore as D, capi
    from microhh_b200.grid import GridData
    from microhh_b200.synthetic import make_case
    it, jt, kt = shape
    z = stretched_z(kt, 2.)
    g = O.Grid(it, jt, kt, 6., 4., 2., 3, 3, 3, dtype, z=z, order=4)
    gd = GridData(it, jt, kt, 6., 4., 2., 3, 3, 3, dtype, z=z, order=4)
    case = make_case(gd, seed=5, noise=0.02)
    ks, ke = g.kstart, g.kend
    case["w"][:ks+1] = 0; case["w"][ke:] = 0
    case["th"] = (1. + 0.1*case["u"]).astype(dtype)            # a passive scalar with sane ghost values
    if jt == 1:
        case["v"][:] = 0
    for n in ("u", "v"):
        case[n + "_bot"] = np.zeros(gd.shape2d, dtype); case[n + "_top"] = np.zeros(gd.shape2d, dtype)
        case[n + "_gradbot"] = np.zeros(gd.shape2d, dtype); case[n + "_gradtop"] = np.zeros(gd.shape2d, dtype)
    case["th_gradbot"] = np.zeros(gd.shape2d, dtype); case["th_gradtop"] = np.zeros(gd.shape2d, dtype)
    ctx = D.Context(gd, 0)
    ones = np.ones(gd.kcells, dtype)
    ctx.set_basestate(ones, ones, 300*ones, 300*ones)
    visc = 1e-3
    f = D.Fields(ctx, case, visc=visc, svisc=visc)
    prm = D.make_params(swadvec=swadvec, swdiff="4", swthermo=None, surface_model=False, mbcbot=mbc, mbctop=mbc)
    oprm = ostep.default_params(); oprm.update(swadvec=swadvec, swdiff="4", visc=visc, svisc=visc, mbcbot=mbc, mbctop=mbc)
    dt = 0.01
    D.Dycore(ctx, prm).step(f, dt)
    ostep.dycore_step(g, O.NumpyKernels(g), case, oprm, dt)
    ctx.sync()
    names = ("u", "w", "th") if jt == 1 else ("u", "v", "w", "th")
    for n in names:
        assert rel_l2(interior(g, f[n].cpu().numpy()), interior(g, case[n])) <= 20*TOL[dtype], n
    un = {c: f[c].cpu().numpy().copy() for c in "uvw"}
    for c in "uvw":
        O.boundary_cyclic(g, un[c])
    O.ghost_cells_w_4th(g, un["w"], True)
    scale = float(np.abs(interior(g, case["u"])).max())/float(g.dx)
    assert float(O.Pres4(g).divergence(un["u"], un["v"], un["w"])) <= (1e-10 if dtype == np.float64 else 5e-3)*scale


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(32, 16, 12), (20, 12, 8), (16, 1, 16)])
def test_advec_s_lim_and_fluxlimit_step(dtype, shape):
    """Koren-limited scalar advection (include/advec_monotonic.h:98-202) for a scalar in fluxlimit_list: the kernel through
    Advec::exec, and a full RK3 step in which the limited scalar bypasses the fused tendency kernel."""
    g, gd, case = make_pair(*shape, dtype, stretched=True, anelastic=True, ns=2)
    rng = np.random.default_rng(9)
    case["s1"] = (case["s1"] + 0.5*rng.standard_normal(gd.shape)).astype(dtype)          # rough enough for every limiter branch
    halo = copy.deepcopy(case); prepare_halos(g, halo)
    from microhh_b200 import dycore as D
    ctx = D.Context(gd, 0)
    ctx.set_basestate(case["rhoref"], case["rhorefh"], case["thref"], case["threfh"])
    f = D.Fields(ctx, halo, scalars=case["scalars"], fluxlimit_list=("s1",))
    D.Advec(ctx, "2i5").exec(f)
    rr, rh = case["rhoref"], case["rhorefh"]
    ref = {n: g.field() for n in ("tht", "s1t")}
    O.advec_2i5_s(g, ref["tht"], halo["th"], halo["u"], halo["v"], halo["w"], rr, rh)
    O.advec_s_lim(g, ref["s1t"], halo["s1"], halo["u"], halo["v"], halo["w"], rr, rh)
    for n in ref:
        assert rel_l2(f[n].cpu().numpy(), ref[n]) <= TOL[dtype], n
    # full step: th fused in the momentum kernel, s1 limited
    f2 = D.Fields(ctx, case, scalars=case["scalars"], fluxlimit_list=("s1",))
    D.Dycore(ctx, D.make_params(ns=2)).step(f2, 2.0)
    oprm = ostep.default_params(); oprm.update(fluxlimit_list=("s1",))
    ostep.dycore_step(g, O.NumpyKernels(g), case, oprm, 2.0)
    ctx.sync()
    for n in ("u", "v", "w", "th", "s1"):
        assert rel_l2(interior(g, f2[n].cpu().numpy()), interior(g, case[n])) <= 5*TOL[dtype], n


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("surface,mason", [(True, True), (True, False), (False, True)])
def test_evisc_neutral_and_neutral_step(dtype, surface, mason):
    """Thermo_type::Disabled: calc_evisc_neutral (src/diff_smag2.cxx:47-146; Mason n = 1 with the surface model, van Driest
    damping over resolved walls) and a full RK3 step without buoyancy."""
    g, gd, case = make_pair(32, 24, 16, dtype, stretched=True, anelastic=True)
    halo = copy.deepcopy(case); prepare_halos(g, halo)
    from microhh_b200 import dycore as D
    ctx = D.Context(gd, 0)
    ctx.set_basestate(case["rhoref"], case["rhorefh"], case["thref"], case["threfh"])
    visc = 1e-2
    f = D.Fields(ctx, halo, visc=visc, svisc=visc)
    prm = D.make_params(swthermo=None, surface_model=surface, sw_mason=mason)
    D.Diff(ctx, prm).exec_viscosity(f)
    ev = g.field()
    O.diff_strain2(g, ev, halo["u"], halo["v"], halo["w"], halo["dudz_mo"], halo["dvdz_mo"], surface)
    O.diff_evisc_neutral(g, ev, halo["u"], halo["v"], halo["w"], halo["z0m"], 0.23, visc, surface, mason)
    k0 = g.kstart - (0 if surface else 1); k1 = g.kend + (0 if surface else 1)
    assert rel_l2(f["evisc"].cpu().numpy()[k0:k1], ev[k0:k1]) <= 10*TOL[dtype]
    if not surface:
        return
    f2 = D.Fields(ctx, case, visc=visc, svisc=visc)
    D.Dycore(ctx, prm).step(f2, 2.0)
    oprm = ostep.default_params(); oprm.update(swthermo=None, surface_model=surface, sw_mason=mason, visc=visc, svisc=visc)
    ostep.dycore_step(g, O.NumpyKernels(g), case, oprm, 2.0)
    ctx.sync()
    for n in ("u", "v", "w", "th"):
        assert rel_l2(interior(g, f2[n].cpu().numpy()), interior(g, case[n])) <= 5*TOL[dtype], n


@pytest.mark.parametrize("dtype", DTYPES)
def test_thermo_dry_buoyancy(dtype):
    g, gd, case = make_pair(32, 16, 12, dtype)
    prepare_halos(g, case)
    D, ctx, f, prm = gpu_setup(gd, case)
    D.Thermo_dry(ctx).exec(f)
    ref = g.field()
    O.thermo_dry_buoyancy_tend_2nd(g, ref, case["th"], case["threfh"])
    assert rel_l2(f["wt"].cpu().numpy(), ref) <= TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("substep", [0, 1, 2])
def test_rk3(dtype, substep):
    g, gd, case = make_pair(16, 12, 8, dtype)
    rng = np.random.default_rng(3)
    for n in ("u", "v", "w", "th"):
        case[n + "t"] = rng.standard_normal(gd.shape).astype(dtype)
    D, ctx, f, prm = gpu_setup(gd, case)
    D.Timeloop(ctx).exec(f, substep, 6.0)
    for n in ("u", "v", "w", "th"):
        a, at = case[n].copy(), case[n + "t"].copy()
        O.rk3(g, a, at, substep, 6.0)
        assert rel_l2(f[n].cpu().numpy(), a) <= TOL[dtype]
        assert rel_l2(f[n + "t"].cpu().numpy(), at) <= TOL[dtype] or np.abs(at).max() == 0
        if substep == 2:
            assert float(f[n + "t"].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(32, 16, 12), (48, 40, 24), (20, 12, 8), (96, 6, 8), (16, 1, 16)])
def test_fft_roundtrip_and_solve(dtype, shape):
    """x/y transforms round-trip to the identity; with the tridiagonal solve in between the result
    equals the oracle's FFTW-semantics solve (r2hc -> tdma -> hc2r)."""
    import torch
    g, gd, case = make_pair(*shape, dtype, stretched=True, anelastic=True)
    D, ctx, f, prm = gpu_setup(gd, case)
    rng = np.random.default_rng(7)
    rhs = rng.standard_normal((gd.kmax, gd.jmax, gd.imax)).astype(dtype)
    a_in = torch.from_numpy(rhs).cuda(); a_out = torch.zeros_like(a_in)
    pres = D.Pres(ctx)
    pres.fft_roundtrip(a_in, a_out, solve=False)
    assert rel_l2(a_out.cpu().numpy(), rhs) <= 20*TOL[dtype]
    pres.fft_roundtrip(a_in, a_out, solve=True)
    P = O.Pres2(g, case["rhoref"], case["rhorefh"])
    p = g.field()
    P.solve(rhs.copy(), p)
    assert rel_l2(a_out.cpu().numpy(), interior(g, p)) <= 50*TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(32, 16, 12), (48, 40, 24)])
@pytest.mark.parametrize("anel", [False, True])
def test_pres_2_exec(dtype, shape, anel):
    g, gd, case = make_pair(*shape, dtype, stretched=True, anelastic=anel)
    prepare_halos(g, case)
    rng = np.random.default_rng(11)
    for n in ("ut", "vt", "wt"):
        t = np.zeros(gd.shape, dtype)
        interior(g, t)[...] = 0.01*rng.standard_normal((gd.kmax, gd.jmax, gd.imax))
        case[n] = t
    case["wt"][:g.kstart+1] = 0
    D, ctx, f, prm = gpu_setup(gd, case)
    sub_dt = 2.0
    D.Pres(ctx).exec(f, sub_dt)
    P = O.Pres2(g, case["rhoref"], case["rhorefh"])
    ref = {n: case[n].copy() for n in ("ut", "vt", "wt")}
    p = g.field()
    P.exec(p, case["u"], case["v"], case["w"], ref["ut"], ref["vt"], ref["wt"], sub_dt)
    ptol = 50*TOL[dtype]
    got_p = f["p"].cpu().numpy()
    assert rel_l2(got_p[g.kstart-1:g.kend], p[g.kstart-1:g.kend]) <= ptol
    for n in ref:
        assert rel_l2(interior(g, f[n].cpu().numpy()), interior(g, ref[n])) <= ptol, n
    # post-pressure divergence of u + dt*ut at the reference's level (both ~ machine epsilon * scale)
    un = {c: case[c] + sub_dt*f[c + "t"].cpu().numpy() for c in "uvw"}
    uo = {c: case[c] + sub_dt*ref[c + "t"] for c in "uvw"}
    for d_ in (un, uo):
        for c in "uvw":
            O.boundary_cyclic(g, d_[c])
    div_gpu = P.divergence(un["u"], un["v"], un["w"]); div_ref = P.divergence(uo["u"], uo["v"], uo["w"])
    scale = np.abs(interior(g, case["u"])).max()/float(g.dx)
    eps = np.finfo(dtype).eps
    assert div_gpu <= max(10*div_ref, 200*eps*scale)
    # check_divergence entry point, on the (divergent) input velocities
    div_in = float(P.divergence(case["u"], case["v"], case["w"]))
    assert abs(D.Pres(ctx).check_divergence(f) - div_in) <= 100*TOL[dtype]*div_in


def run_steps(dtype, shape, nsteps, anel, stretched, ns=1):
    g, gd, case = make_pair(*shape, dtype, stretched=stretched, anelastic=anel, ns=ns)
    D, ctx, f, _ = gpu_setup(gd, case, ns)
    prm = D.make_params(ns=ns)
    dyc = D.Dycore(ctx, prm)
    oprm = ostep.default_params()
    K = O.NumpyKernels(g)
    dt = 2.0
    for _ in range(nsteps):
        dyc.step(f, dt)
        ostep.dycore_step(g, K, case, oprm, dt)
    ctx.sync()
    return g, case, f


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,anel,stretched", [((32, 32, 32), False, False), ((48, 24, 16), True, True)])
def test_full_rk3_step(dtype, shape, anel, stretched):
    """One full RK3 step (3 sub-steps) of the fused path == the reference call order on the oracle:
    relative L2 on u, v, w, th within BASELINE.json's tolerance."""
    g, case, f = run_steps(dtype, shape, 1, anel, stretched)
    for n in ("u", "v", "w", "th"):
        assert rel_l2(interior(g, f[n].cpu().numpy()), interior(g, case[n])) <= TOL[dtype], n
    for n in ("ut", "vt", "wt", "tht"):
        assert float(f[n].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", DTYPES)
def test_two_steps_two_scalars(dtype):
    g, case, f = run_steps(dtype, (32, 16, 16), 2, True, True, ns=2)
    for n in ("u", "v", "w", "th", "s1"):
        assert rel_l2(interior(g, f[n].cpu().numpy()), interior(g, case[n])) <= 5*TOL[dtype], n


def test_errors_are_reported():
    import ctypes as C
    g, gd, case = make_pair(16, 12, 8, np.float64)
    D, ctx, f, prm = gpu_setup(gd, case)
    rc = ctx.lib.mhh_boundary_cyclic(ctx.h, None, 2)
    assert rc == -1 and b"NULL" in ctx.lib.mhh_last_error(ctx.h)
    with pytest.raises(D.MhhError):
        D.Advec(ctx, "4").exec(f)          # 4th-order scheme on a 2nd-order grid
