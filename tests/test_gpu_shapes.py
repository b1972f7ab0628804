"""
Parity on the shapes BASELINE.json names and on the kernel instantiations bench.py times:
one full RK3 step through the C ABI against the oracle driven by the reference's own compiled CPU kernels
(oracle/_ref/libmhh_ref.so; numpy restatement when the library is absent) at

  * drycblles 128^3 fp64 with swadvec=2i5 and as shipped (swadvec=2 + smag2)  -- cases/drycblles/drycblles.ini:6-25
  * moser180-shaped 256 x 192 x 128 4th-order DNS (advec_4 + diff_4 + pres_4) -- cases/moser180/moser180.ini
  * 256 x 256 x 128 fp32 (USESP)
  * multi-x-tile grids with a ragged last tile (itot = 160, 192), rows not a multiple of the CTA height
  * every warp-FFT length the library instantiates (x: itot/2 = 16..1024, y: jtot = 8..2048): round trip and solve

Tolerances are BASELINE.json's: relative L2 <= 1e-12 (fp64) / 1e-5 (fp32) on u, v, w, th.
"""
import numpy as np
import pytest

from util import TOL, rel_l2, make_pair, interior, prepare_halos, stretched_z
from oracle import oracle as O
from oracle import step as ostep
from oracle import refbind

pytestmark = pytest.mark.gpu


def kernels(g):
    return refbind.RefKernels(g, fast=False) if refbind.available(False) else O.NumpyKernels(g)


def gpu_setup(gd, case, ns=1, **fkw):
    from microhh_b200 import dycore as D
    ctx = D.Context(gd, 0)
    ctx.set_basestate(case["rhoref"], case["rhorefh"], case["thref"], case["threfh"])
    f = D.Fields(ctx, case, scalars=case["scalars"], **fkw)
    return D, ctx, f


def check_step(g, f, case, names, tol):
    errs = {}
    for n in names:
        errs[n] = rel_l2(interior(g, f[n].cpu().numpy()), interior(g, case[n]))
    assert all(e <= tol for e in errs.values()), errs
    return errs


def truth_fp64(case32, shape, dt, oprm, **pair_kw):
    """The same step in fp64 on the same (fp32-representable) inputs: the yardstick for fp32 runs on grids where two
    independent fp32 evaluations differ by more than 1e-5 from each other through rounding alone (the Poisson solve
    amplifies the rounding noise of the divergence with the grid size)."""
    import copy
    g64, gd64, c64 = make_pair(*shape, np.float64, **pair_kw)
    for k, a in case32.items():
        if isinstance(a, np.ndarray):
            c64[k] = a.astype(np.float64)
    ostep.dycore_step(g64, kernels(g64), c64, copy.deepcopy(oprm), dt)
    return g64, c64


def check_fp32_at_reference_level(g, f, case32, g64, c64, names):
    """fp32 on large grids: the GPU result must be as close to the fp64 truth as the reference's own fp32 CPU kernels are
    (<= 1.25 x the reference's error + 1e-6), which is the strongest statement rounding allows."""
    out = {}
    for n in names:
        t = interior(g64, c64[n])
        e_gpu = rel_l2(interior(g, f[n].cpu().numpy()), t)
        e_ref = rel_l2(interior(g, case32[n]), t)
        out[n] = (e_gpu, e_ref)
    assert all(eg <= 1.25*er + 1e-6 for eg, er in out.values()), out
    return out


# ------------------------------------------------------------------------------------------------ LES configurations
@pytest.mark.parametrize("swadvec,igc", [("2i5", 3), ("2i5", 4), ("2", 3), ("2", 4)])
def test_drycblles_128_fp64(swadvec, igc):
    """configs[0]: drycblles 128^3 fp64 (3200 m cube, dt = 6 s is the case's dtmax); `2` is the .ini as shipped; igc = 4 is what
    the adapters ask Grid for (aligned pairs in the TMA-staged kernel), igc = 3 the reference's minimum for advec_2i5."""
    g, gd, case = make_pair(128, 128, 128, np.float64, igc=igc)
    D, ctx, f = gpu_setup(gd, case)
    prm = D.make_params(swadvec=swadvec)
    oprm = ostep.default_params(); oprm.update(swadvec=swadvec)
    dt = 6.0
    D.Dycore(ctx, prm).step(f, dt)
    ostep.dycore_step(g, kernels(g), case, oprm, dt)
    ctx.sync()
    check_step(g, f, case, ("u", "v", "w", "th"), TOL[np.float64])
    # post-pressure divergence at the reference's level
    P = O.Pres2(g, case["rhoref"], case["rhorefh"])
    un = {c: f[c].cpu().numpy().copy() for c in "uvw"}
    for c in "uvw":
        O.boundary_cyclic(g, un[c]); O.boundary_cyclic(g, case[c])
    div_gpu = float(P.divergence(un["u"], un["v"], un["w"])); div_ref = float(P.divergence(case["u"], case["v"], case["w"]))
    scale = float(np.abs(interior(g, case["u"])).max())/float(g.dx)
    assert div_gpu <= max(10*div_ref, 200*np.finfo(np.float64).eps*scale), (div_gpu, div_ref)


@pytest.mark.parametrize("dtype,igc", [(np.float64, 3), (np.float64, 4), (np.float32, 4)])
@pytest.mark.parametrize("shape", [(160, 20, 24), (128, 10, 33)])
@pytest.mark.parametrize("swthermo,surface", [("dry", True), (None, True), ("dry", False)])
def test_advec2_smag2_fused(dtype, igc, shape, swthermo, surface, monkeypatch):
    """cases/drycblles as shipped (swadvec = 2 + smag2): Advec_2's fluxes run inside the TMA-staged fused tendency kernel
    (mom3_kernel<..., ADV2>) when the grid has igc >= 3; multi-tile / ragged grids, with and without thermo and surface model.
    The same step with MHH_FUSE_ADVEC2=0 (Advec_2 point-wise, then the diffusion alone) must agree with it too."""
    kw = dict(stretched=True, anelastic=True, igc=igc)
    g, gd, case = make_pair(*shape, dtype, **kw)
    oprm = ostep.default_params(); oprm.update(swadvec="2", swthermo=swthermo, surface_model=surface, visc=1e-2 if not surface else 1e-5)
    dt = 2.0
    names = ["u", "v", "w", "th"]
    if dtype == np.float32:
        g64, c64 = truth_fp64(case, shape, dt, oprm, **kw)
    res = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("MHH_FUSE_ADVEC2", fused)
        D, ctx, f = gpu_setup(gd, case, visc=oprm["visc"])
        prm = D.make_params(swadvec="2", swthermo=swthermo or "0", surface_model=surface)
        ctx.profile_start()
        D.Dycore(ctx, prm).step(f, dt)
        prof = ctx.profile_stop()
        ctx.sync()
        assert ("mom3_kernel_advec2" in prof) == (fused == "1"), sorted(prof)
        res[fused] = {n: f[n].cpu().numpy() for n in names}
        if fused == "1":
            f1 = f
    ostep.dycore_step(g, kernels(g), case, oprm, dt)
    if dtype == np.float32:
        check_fp32_at_reference_level(g, f1, case, g64, c64, names)
        check_step(g, f1, case, names, 5*TOL[dtype])
    else:
        check_step(g, f1, case, names, TOL[dtype])
    for n in names:
        assert rel_l2(interior(g, res["1"][n]), interior(g, res["0"][n])) <= (TOL[dtype] if dtype == np.float64 else 5*TOL[dtype]), n


@pytest.mark.parametrize("igc", [3, 4])
def test_les_256x256x128_fp32(igc):
    """USESP build of the LES path (the bomex-shaped grid is 512 x 512 x 256; this is its 1/8 at oracle-friendly cost).
    igc = 4 (16-byte row pitch) takes the TMA-staged kernels, igc = 3 the cp.async ones."""
    kw = dict(stretched=True, anelastic=True, sizes=(6400., 6400., 3200.), igc=igc)
    g, gd, case = make_pair(256, 256, 128, np.float32, **kw)
    D, ctx, f = gpu_setup(gd, case)
    dt = 4.0
    g64, c64 = truth_fp64(case, (256, 256, 128), dt, ostep.default_params(), **kw)
    D.Dycore(ctx, D.make_params()).step(f, dt)
    ostep.dycore_step(g, kernels(g), case, ostep.default_params(), dt)
    ctx.sync()
    check_fp32_at_reference_level(g, f, case, g64, c64, ("u", "v", "w", "th"))


@pytest.mark.parametrize("dtype,igc", [(np.float64, 3), (np.float32, 3), (np.float32, 4), (np.float64, 4)])
@pytest.mark.parametrize("shape", [(160, 20, 24), (192, 36, 40), (128, 10, 33), (64, 9, 17)])
@pytest.mark.parametrize("ns", [1, 2])
def test_multi_tile_step(dtype, igc, shape, ns):
    """>= 2 x-tiles of the 64-wide marching kernels (blockIdx.x > 0 TMA coordinates), ragged last x tile (160 = 2.5
    tiles), rows that are not a multiple of the CTA height, ktot that splits into uneven z-chunks; one and two scalars."""
    kw = dict(stretched=True, anelastic=True, ns=ns, igc=igc)
    g, gd, case = make_pair(*shape, dtype, **kw)
    D, ctx, f = gpu_setup(gd, case, ns)
    dt = 2.0
    names = ["u", "v", "w"] + case["scalars"]
    if dtype == np.float32:
        g64, c64 = truth_fp64(case, shape, dt, ostep.default_params(), **kw)
    D.Dycore(ctx, D.make_params(ns=ns)).step(f, dt)
    ostep.dycore_step(g, kernels(g), case, ostep.default_params(), dt)
    ctx.sync()
    if dtype == np.float32:
        check_fp32_at_reference_level(g, f, case, g64, c64, names)
        check_step(g, f, case, names, 5*TOL[dtype])          # and never far from the reference's fp32 result itself
    else:
        check_step(g, f, case, names, TOL[dtype])


@pytest.mark.parametrize("dtype,igc", [(np.float64, 3), (np.float64, 4), (np.float32, 3), (np.float32, 4), (np.float32, 6)])
@pytest.mark.parametrize("itot", [128, 160, 192])
@pytest.mark.parametrize("surface", [True, False])
def test_fused_tendencies_multi_tile(dtype, igc, itot, surface):
    """The fused tendency stage alone (thermo.exec + advec.exec + diff.exec = mom3 / tile kernels) and the eddy-viscosity
    tile kernel on multi-tile grids, against the individual oracle kernels."""
    g, gd, case = make_pair(itot, 25, 20, dtype, stretched=True, anelastic=True, igc=igc)
    prepare_halos(g, case)
    D, ctx, f = gpu_setup(gd, case)
    prm = D.make_params(surface_model=surface)
    K = kernels(g)
    D.Diff(ctx, prm).exec_viscosity(f)
    ev = g.field(); N2 = g.field()
    K.diff_strain2(ev, case["u"], case["v"], case["w"], case["dudz_mo"], case["dvdz_mo"], surface)
    K.thermo_dry_N2(N2, case["th"], case["thref"])
    K.diff_evisc(ev, case["u"], case["v"], case["w"], N2, case["dbdz_mo"], case["z0m"], 0.23, 1./3., surface, True)
    K.boundary_cyclic(ev)
    k0 = g.kstart - (0 if surface else 1); k1 = g.kend + (0 if surface else 1)
    assert rel_l2(f["evisc"].cpu().numpy()[k0:k1], ev[k0:k1]) <= TOL[dtype]
    import torch
    f["evisc"].copy_(torch.from_numpy(ev))
    D.Dycore(ctx, prm).tendencies(f)
    rr, rh = case["rhoref"], case["rhorefh"]
    ref = {n: g.field() for n in ("ut", "vt", "wt", "tht")}
    K.thermo_dry_buoyancy_tend_2nd(ref["wt"], case["th"], case["threfh"])
    K.advec_2i5_u(ref["ut"], case["u"], case["v"], case["w"], rr, rh)
    K.advec_2i5_v(ref["vt"], case["u"], case["v"], case["w"], rr, rh)
    K.advec_2i5_w(ref["wt"], case["u"], case["v"], case["w"], rr, rh)
    K.advec_2i5_s(ref["tht"], case["th"], case["u"], case["v"], case["w"], rr, rh)
    K.diff_u(ref["ut"], case["u"], case["v"], case["w"], ev, case["u_fluxbot"], case["u_fluxtop"], rr, rh, 1e-5, surface)
    K.diff_v(ref["vt"], case["u"], case["v"], case["w"], ev, case["v_fluxbot"], case["v_fluxtop"], rr, rh, 1e-5, surface)
    K.diff_w(ref["wt"], case["u"], case["v"], case["w"], ev, rr, rh, 1e-5)
    K.diff_c(ref["tht"], case["th"], ev, case["th_fluxbot"], case["th_fluxtop"], rr, rh, 1./3., 1e-5, surface)
    for n in ref:
        # advection and diffusion are summed in a different order than the reference's two passes: a few ulp
        assert rel_l2(interior(g, f[n].cpu().numpy()), interior(g, ref[n])) <= 5*TOL[dtype], n


# ------------------------------------------------------------------------------------------------ 4th-order DNS
@pytest.mark.parametrize("swadvec", ["4m", "4"])
def test_moser180_shape_order4_fp64(swadvec):
    """configs[2]: moser180-shaped 256 x 192 x 128 channel (no-slip walls, stretched z): advec_4m + diff_4 + pres_4 as
    cases/moser180/moser180.ini ships it, and the plain advec_4 variant."""
    from microhh_b200.grid import GridData
    from microhh_b200.synthetic import make_case
    from microhh_b200 import dycore as D
    dtype = np.float64
    it, jt, kt = 256, 192, 128
    z = stretched_z(kt, 2.)
    sizes = (2*np.pi, np.pi, 2.)          # cases/moser180/moser180.ini: xsize, ysize, zsize
    g = O.Grid(it, jt, kt, *sizes, 3, 3, 3, dtype, z=z, order=4)
    gd = GridData(it, jt, kt, *sizes, 3, 3, 3, dtype, z=z, order=4)
    case = make_case(gd, seed=5, noise=0.02)
    ks, ke = g.kstart, g.kend
    case["w"][:ks+1] = 0; case["w"][ke:] = 0
    case["th"] = (1. + 0.1*case["u"]).astype(dtype)
    for n in ("u", "v"):
        for s in ("_bot", "_top", "_gradbot", "_gradtop"):
            case[n + s] = np.zeros(gd.shape2d, dtype)
    case["th_gradbot"] = np.zeros(gd.shape2d, dtype); case["th_gradtop"] = np.zeros(gd.shape2d, dtype)
    ctx = D.Context(gd, 0)
    ones = np.ones(gd.kcells, dtype)
    ctx.set_basestate(ones, ones, 300*ones, 300*ones)
    visc = 1e-3
    f = D.Fields(ctx, case, visc=visc, svisc=visc)
    prm = D.make_params(swadvec=swadvec, swdiff="4", swthermo=None, surface_model=False, mbcbot=0, mbctop=0)
    oprm = ostep.default_params(); oprm.update(swadvec=swadvec, swdiff="4", visc=visc, svisc=visc, mbcbot=0, mbctop=0,
                                               swthermo=None, surface_model=False)
    dt = 0.002
    D.Dycore(ctx, prm).step(f, dt)
    ostep.dycore_step(g, kernels(g), case, oprm, dt)
    ctx.sync()
    check_step(g, f, case, ("u", "v", "w", "th"), TOL[dtype])
    un = {c: f[c].cpu().numpy().copy() for c in "uvw"}
    for c in "uvw":
        O.boundary_cyclic(g, un[c])
    O.ghost_cells_w_4th(g, un["w"], True)
    scale = float(np.abs(interior(g, case["u"])).max())/float(g.dx)
    assert float(O.Pres4(g).divergence(un["u"], un["v"], un["w"])) <= 1e-10*scale


# ------------------------------------------------------------------------------------------------ every FFT length
X_LENGTHS = [32, 64, 128, 256, 512, 1024, 2048]          # itot: the warp kernel is instantiated for itot/2 = 16 .. 1024
Y_LENGTHS = [8, 16, 32, 64, 128, 256, 512, 1024, 2048]   # jtot


def _fft_case(itot, jtot, ktot, dtype):
    g, gd, case = make_pair(itot, jtot, ktot, dtype, stretched=True, anelastic=True)
    D, ctx, f = gpu_setup(gd, case)
    import torch
    rng = np.random.default_rng(itot + 7*jtot)
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


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("itot", X_LENGTHS)
def test_wfft_x_lengths(dtype, itot):
    if dtype == np.float64 and itot == 2048:
        # fp64 rows of 1024 complex points: 8 warp rows still fit 227 KB (wfft_fits), keep it covered
        pass
    _fft_case(itot, 8, 6, dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("jtot", Y_LENGTHS)
def test_wfft_y_lengths(dtype, jtot):
    _fft_case(32, jtot, 6, dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(96, 40, 8), (60, 18, 6), (80, 50, 7), (24, 30, 6)])
def test_generic_fft_lengths(dtype, shape):
    """2^a 3^b 5^c lengths go through the generic mixed-radix block kernels."""
    _fft_case(*shape, dtype)
