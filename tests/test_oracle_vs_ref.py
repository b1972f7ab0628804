"""
Pins the oracle (CPU, no GPU needed): the numpy restatement in oracle/oracle.py against the
reference's own CPU kernels compiled from /root/reference into oracle/_ref/libmhh_ref.so
(-O2 -ffp-contract=off).  Bit-exact for every kernel the harness reaches, and for a full RK3
step assembled in the reference's call order.  Skipped (not failed) where oracle/_ref has not
been built; the committed golden vectors (tests/golden/, test_golden.py) cover that case.
"""
import copy
import numpy as np
import pytest

from util import moist_case, make_moist_pair, make_pair, prepare_halos, interior, add_sgstke
from oracle import oracle as O
from oracle import step as ostep
from oracle import refbind

pytestmark = pytest.mark.skipif(not refbind.available(), reason="oracle/_ref/libmhh_ref.so not built (make -C oracle)")

DTYPES = [np.float64, np.float32]
CASES = [((16, 12, 8), False, False), ((20, 12, 10), True, True), ((24, 1, 8), False, True)]


def both(g):
    return O.NumpyKernels(g), refbind.RefKernels(g)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,anel,stretched", CASES)
def test_kernels_bitexact(dtype, shape, anel, stretched):
    g, gd, case = make_pair(*shape, dtype, stretched=stretched, anelastic=anel)
    prepare_halos(g, case)
    N, R = both(g)
    rr, rh = case["rhoref"], case["rhorefh"]
    rng = np.random.default_rng(11)

    def pair(fn_name, out_names, *args_builder):
        outs = []
        for K in (N, R):
            c = copy.deepcopy(case)
            for n in out_names:
                c[n] = rng_fill[n].copy()
            args_builder[0](K, c)
            outs.append([c[n] for n in out_names])
        for a, b, n in zip(outs[0], outs[1], out_names):
            assert np.array_equal(a, b), (fn_name, n, float(np.abs(a.astype(np.float64) - b).max()))

    rng_fill = {n: rng.standard_normal(gd.shape).astype(dtype) for n in ("ut", "vt", "wt", "tht", "evisc")}

    pair("advec_u", ["ut"], lambda K, c: K.advec_2i5_u(c["ut"], c["u"], c["v"], c["w"], rr, rh))
    pair("advec_v", ["vt"], lambda K, c: K.advec_2i5_v(c["vt"], c["u"], c["v"], c["w"], rr, rh))
    pair("advec_w", ["wt"], lambda K, c: K.advec_2i5_w(c["wt"], c["u"], c["v"], c["w"], rr, rh))
    pair("advec_s", ["tht"], lambda K, c: K.advec_2i5_s(c["tht"], c["th"], c["u"], c["v"], c["w"], rr, rh))
    pair("advec_s_lim", ["tht"], lambda K, c: K.advec_s_lim(c["tht"], c["th"], c["u"], c["v"], c["w"], rr, rh))
    rough = np.random.default_rng(5).standard_normal(gd.shape).astype(dtype)       # every branch of the limiter
    rough[:, ::3, ::2] = rough[:, 1::3, ::2][:, :rough[:, ::3, ::2].shape[1]] if shape[1] > 3 else 0.   # and exact ties (denominator guard)
    pair("advec_s_lim_rough", ["tht"], lambda K, c: K.advec_s_lim(c["tht"], rough, c["u"], c["v"], c["w"], rr, rh))
    # Advec_2 / Diff_2 (reference src/advec_2.cxx, src/diff_2.cxx)
    pair("advec_2_u", ["ut"], lambda K, c: K.advec_2_u(c["ut"], c["u"], c["v"], c["w"], rr, rh))
    pair("advec_2_v", ["vt"], lambda K, c: K.advec_2_v(c["vt"], c["u"], c["v"], c["w"], rr, rh))
    pair("advec_2_w", ["wt"], lambda K, c: K.advec_2_w(c["wt"], c["u"], c["v"], c["w"], rr, rh))
    pair("advec_2_s", ["tht"], lambda K, c: K.advec_2_s(c["tht"], c["th"], c["u"], c["v"], c["w"], rr, rh))
    pair("diff_2_c", ["ut"], lambda K, c: K.diff_2_c(c["ut"], c["u"], 0.7))
    pair("diff_2_c_s", ["tht"], lambda K, c: K.diff_2_c(c["tht"], c["th"], 1.3))
    pair("diff_2_w", ["wt"], lambda K, c: K.diff_2_w(c["wt"], c["w"], 0.7))
    assert N.advec_2_cfl(case["u"], case["v"], case["w"], 2.0) == R.advec_2_cfl(case["u"], case["v"], case["w"], 2.0)
    pair("buoyancy", ["wt"], lambda K, c: K.thermo_dry_buoyancy_tend_2nd(c["wt"], c["th"], c["threfh"]))

    for surface in (True, False):
        # eddy viscosity chain: strain2 -> N2 -> evisc
        ev = []
        for K in (N, R):
            c = copy.deepcopy(case)
            c["evisc"] = np.zeros(gd.shape, dtype)
            K.diff_strain2(c["evisc"], c["u"], c["v"], c["w"], c["dudz_mo"], c["dvdz_mo"], surface)
            n2 = np.zeros(gd.shape, dtype)
            K.thermo_dry_N2(n2, c["th"], c["thref"])
            K.diff_evisc(c["evisc"], c["u"], c["v"], c["w"], n2, c["dbdz_mo"], c["z0m"], 0.23, 1./3., surface, True)
            ev.append(c["evisc"])
        assert np.array_equal(ev[0], ev[1]), ("evisc", surface)
        # neutral variant (no thermo): Mason with n = 1 (surface model) / van Driest damping (resolved walls)
        for mason in (True, False):
            evn = []
            for K in (N, R):
                c = copy.deepcopy(case)
                c["evisc"] = np.zeros(gd.shape, dtype)
                K.diff_strain2(c["evisc"], c["u"], c["v"], c["w"], c["dudz_mo"], c["dvdz_mo"], surface)
                K.diff_evisc_neutral(c["evisc"], c["u"], c["v"], c["w"], c["z0m"], 0.23, 1e-2, surface, mason)
                evn.append(c["evisc"])
            assert np.array_equal(evn[0], evn[1]), ("evisc_neutral", surface, mason)
            assert not np.array_equal(evn[0], ev[0])
        case_e = copy.deepcopy(case); case_e["evisc"] = ev[1]
        for nm, fn in (("diff_u", lambda K, c: K.diff_u(c["ut"], c["u"], c["v"], c["w"], c["evisc"], c["u_fluxbot"], c["u_fluxtop"], rr, rh, 1e-5, surface)),
                       ("diff_v", lambda K, c: K.diff_v(c["vt"], c["u"], c["v"], c["w"], c["evisc"], c["v_fluxbot"], c["v_fluxtop"], rr, rh, 1e-5, surface)),
                       ("diff_w", lambda K, c: K.diff_w(c["wt"], c["u"], c["v"], c["w"], c["evisc"], rr, rh, 1e-5)),
                       ("diff_c", lambda K, c: K.diff_c(c["tht"], c["th"], c["evisc"], c["th_fluxbot"], c["th_fluxtop"], rr, rh, 1./3., 1e-5, surface))):
            res = []
            for K in (N, R):
                c = copy.deepcopy(case_e)
                for n in ("ut", "vt", "wt", "tht"):
                    c[n] = rng_fill[n].copy()
                fn(K, c)
                res.append([c[n] for n in ("ut", "vt", "wt", "tht")])
            for a, b in zip(*res):
                assert np.array_equal(a, b), (nm, surface)
        assert N.diff_dnmul(ev[0], 1./3.) == R.diff_dnmul(ev[1], 1./3.)
    assert N.advec_2i5_cfl(case["u"], case["v"], case["w"], 2.0) == R.advec_2i5_cfl(case["u"], case["v"], case["w"], 2.0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("edge", [0, 1, 2])
@pytest.mark.parametrize("shape", [(16, 12, 8), (24, 1, 8)])
def test_boundary_cyclic_bitexact(dtype, edge, shape):
    g, gd, case = make_pair(*shape, dtype)
    a = np.random.default_rng(3).standard_normal(gd.shape).astype(dtype)
    b = a.copy()
    N, R = both(g)
    N.boundary_cyclic(a, edge); R.boundary_cyclic(b, edge)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("substep", [0, 1, 2])
def test_rk3_and_tdma_bitexact(dtype, substep):
    g, gd, case = make_pair(16, 12, 8, dtype, stretched=True, anelastic=True)
    N, R = both(g)
    rng = np.random.default_rng(4)
    a0 = rng.standard_normal(gd.shape).astype(dtype); t0 = rng.standard_normal(gd.shape).astype(dtype)
    a1, t1 = a0.copy(), t0.copy()
    N.rk3(a0, t0, substep, 1.7); R.rk3(a1, t1, substep, 1.7)
    assert np.array_equal(a0, a1) and np.array_equal(t0, t1)
    # tridiagonal solve on the spectral right-hand side of this grid
    P = O.Pres2(g, case["rhoref"], case["rhorefh"])
    p = rng.standard_normal((g.kmax, g.jmax, g.imax)).astype(dtype)
    pc0 = np.zeros(gd.shape, dtype); pc1 = np.zeros(gd.shape, dtype)
    q0 = p.copy(); q1 = p.copy()
    P.solve(q0, pc0)
    P.solve(q1, pc1, tdma=lambda pp, b: R.tdma(P.a.copy(), b, P.c.copy(), pp))
    assert np.array_equal(pc0, pc1)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(16, 12, 10), (24, 1, 8), (12, 10, 6)])
@pytest.mark.parametrize("stretched", [False, True])
def test_order4_kernels_bitexact(dtype, shape, stretched):
    """Advec_4 (src/advec_4.cxx) and Diff_4 (src/diff_4.cxx), incl. the 2-D (jtot = 1) variants: numpy oracle == compiled
    reference, bit for bit, on a 4th-order grid (three ghost cells in every direction)."""
    from util import stretched_z
    rng = np.random.default_rng(3)
    it, jt, kt = shape
    g = O.Grid(it, jt, kt, 3200., 3200., 3200., 3, 3, 3, dtype, z=stretched_z(kt, 3200.) if stretched else None, order=4)
    N, R = both(g)
    fld = lambda: rng.standard_normal((g.kcells, g.jcells, g.icells)).astype(dtype)
    u, v, w, s = fld(), fld(), fld(), fld()
    t0 = {n: fld() for n in ("ut", "vt", "wt", "st")}
    res = []
    for K in (N, R):
        t = {n: a.copy() for n, a in t0.items()}
        K.advec_4_u(t["ut"], u, v, w); K.advec_4_v(t["vt"], u, v, w); K.advec_4_w(t["wt"], u, v, w)
        K.advec_4_s(t["st"], s, u, v, w)
        K.diff_4_c(t["ut"], u, 0.7); K.diff_4_c(t["vt"], v, 0.7); K.diff_4_w(t["wt"], w, 0.7); K.diff_4_c(t["st"], s, 1.3)
        res.append((t, K.advec_4_cfl(u, v, w, 2.0)))
    for n in t0:
        assert np.array_equal(res[0][0][n], res[1][0][n]), (n, float(np.abs(res[0][0][n].astype(np.float64) - res[1][0][n]).max()))
        assert not np.array_equal(res[0][0][n], t0[n])
    assert res[0][1] == res[1][1]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(16, 12, 10), (24, 1, 8), (12, 10, 6)])
@pytest.mark.parametrize("stretched", [False, True])
def test_advec_4m_bitexact(dtype, shape, stretched):
    """Advec_4m (src/advec_4m.cxx:51-415, the fully conservative 4th-order scheme moser180 ships with): numpy oracle ==
    compiled reference, bit for bit, tendencies and CFL number; it is not Advec_4 under another name."""
    from util import stretched_z
    rng = np.random.default_rng(5)
    it, jt, kt = shape
    g = O.Grid(it, jt, kt, 6.28, 3.14, 2., 3, 3, 3, dtype, z=stretched_z(kt, 2.) if stretched else None, order=4)
    N, R = both(g)
    fld = lambda: rng.standard_normal((g.kcells, g.jcells, g.icells)).astype(dtype)
    u, v, w, s = fld(), fld(), fld(), fld()
    t0 = {n: fld() for n in ("ut", "vt", "wt", "st")}
    res = []
    for K in (N, R):
        t = {n: a.copy() for n, a in t0.items()}
        K.advec_4m_u(t["ut"], u, v, w); K.advec_4m_v(t["vt"], u, v, w); K.advec_4m_w(t["wt"], u, v, w)
        K.advec_4m_s(t["st"], s, u, v, w)
        res.append((t, K.advec_4m_cfl(u, v, w, 2.0)))
    for n in t0:
        assert np.array_equal(res[0][0][n], res[1][0][n]), (n, float(np.abs(res[0][0][n].astype(np.float64) - res[1][0][n]).max()))
        assert not np.array_equal(res[0][0][n], t0[n])
    assert res[0][1] == res[1][1]
    t4 = t0["ut"].copy(); N.advec_4_u(t4, u, v, w)
    assert not np.array_equal(t4, res[0][0]["ut"])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("surface", [True, False])
def test_full_rk3_step_neutral_bitexact(dtype, surface):
    """LES step without thermo (swthermo=0): calc_evisc_neutral, no buoyancy."""
    g, gd, case = make_pair(20, 12, 10, dtype, stretched=True, anelastic=True)
    c0, c1 = copy.deepcopy(case), copy.deepcopy(case)
    N, R = both(g)
    prm = ostep.default_params(); prm.update(swthermo=None, surface_model=surface, visc=1e-2)
    ostep.dycore_step(g, N, c0, prm, 2.0)
    ostep.dycore_step(g, R, c1, prm, 2.0, pres=refbind.RefPres(g, 2, c1["rhoref"], c1["rhorefh"]))    # the reference's own Pres_2
    for n in ("u", "v", "w", "th", "evisc"):
        assert np.array_equal(c0[n], c1[n]), n


@pytest.mark.parametrize("dtype", DTYPES)
def test_full_rk3_step_fluxlimit_bitexact(dtype):
    """LES step with the second scalar in [advec] fluxlimit_list (Koren-limited advection, src/advec_2i5.cxx:1046-1056)."""
    g, gd, case = make_pair(20, 12, 10, dtype, stretched=True, anelastic=True, ns=2)
    c0, c1 = copy.deepcopy(case), copy.deepcopy(case)
    N, R = both(g)
    prm = ostep.default_params(); prm.update(fluxlimit_list=("s1",))
    ostep.dycore_step(g, N, c0, prm, 2.0)
    ostep.dycore_step(g, R, c1, prm, 2.0, pres=refbind.RefPres(g, 2, c1["rhoref"], c1["rhorefh"]))    # the reference's own Pres_2
    for n in ("u", "v", "w", "th", "s1"):
        assert np.array_equal(c0[n], c1[n]), n
    c2 = copy.deepcopy(case)
    ostep.dycore_step(g, N, c2, ostep.default_params(), 2.0)
    assert not np.array_equal(c0["s1"], c2["s1"])          # the limiter does change the scalar


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(16, 12, 10), (24, 1, 8)])
@pytest.mark.parametrize("swadvec", ["4", "4m"])
def test_full_rk3_step_order4_bitexact(dtype, shape, swadvec):
    """4th-order DNS step (advec_4 or advec_4m + diff_4 + pres_4 + 4th-order ghost cells) in Model::exec order: the numpy
    oracle == the compiled reference (its kernels AND its own Pres_4 / FFT / Grid member functions)."""
    from util import stretched_z
    from microhh_b200.grid import GridData
    from microhh_b200.synthetic import make_case
    it, jt, kt = shape
    z = stretched_z(kt, 2.)
    g = O.Grid(it, jt, kt, 6., 4., 2., 3, 3, 3, dtype, z=z, order=4)
    gd = GridData(it, jt, kt, 6., 4., 2., 3, 3, 3, dtype, z=z, order=4)
    case = make_case(gd, seed=5, noise=0.02)
    ks, ke = g.kstart, g.kend
    case["w"][:ks+1] = 0; case["w"][ke:] = 0
    case["th"] = (1. + 0.1*case["u"]).astype(dtype)
    for n in ("u", "v"):
        for sfx in ("_bot", "_top", "_gradbot", "_gradtop"):
            case[n + sfx] = np.zeros(gd.shape2d, dtype)
    case["th_gradbot"] = np.zeros(gd.shape2d, dtype); case["th_gradtop"] = np.zeros(gd.shape2d, dtype)
    prm = ostep.default_params(); prm.update(swadvec=swadvec, swdiff="4", visc=1e-3, svisc=1e-3, mbcbot=0, mbctop=0)
    c0, c1 = copy.deepcopy(case), copy.deepcopy(case)
    N, R = both(g)
    ostep.dycore_step(g, N, c0, prm, 0.01)
    ostep.dycore_step(g, R, c1, prm, 0.01, pres=refbind.RefPres(g, 4))                                  # the reference's own Pres_4
    for n in ("u", "v", "w", "th", "p"):
        assert np.array_equal(interior(g, c0[n]), interior(g, c1[n])), n
    assert np.isfinite(interior(g, c0["u"])).all() and not np.array_equal(c0["u"], case["u"])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("bc", [0, 1])
def test_order4_ghost_cells_bitexact(dtype, bc):
    """4th-order vertical ghost cells (src/boundary.cxx:776-922): numpy oracle == compiled reference."""
    from util import stretched_z
    rng = np.random.default_rng(8)
    g = O.Grid(12, 10, 8, 3200., 3200., 3200., 3, 3, 3, dtype, z=stretched_z(8, 3200.), order=4)
    N, R = both(g)
    a0 = rng.standard_normal((g.kcells, g.jcells, g.icells)).astype(dtype)
    two = lambda: rng.standard_normal((g.jcells, g.icells)).astype(dtype)
    bot, gbot, top, gtop = two(), two(), two(), two()
    out = []
    for K in (N, R):
        a = a0.copy(); w1 = a0.copy(); w2 = a0.copy()
        K.ghost_cells_bot_4th(a, bc, bot, gbot); K.ghost_cells_top_4th(a, bc, top, gtop)
        K.ghost_cells_w_4th(w1, False); K.ghost_cells_w_4th(w2, True)
        out.append((a, w1, w2))
    for x, y in zip(*out):
        assert np.array_equal(x, y)
    assert not np.array_equal(out[0][0], a0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("swadvec,swdiff", [("2", "2"), ("2", "smag2"), ("2i5", "2")])
def test_full_rk3_step_scheme_combinations_bitexact(dtype, swadvec, swdiff):
    """Advec_2 / Diff_2 in every combination with the LES schemes ("2" + "smag2" is drycblles as shipped)."""
    g, gd, case = make_pair(20, 12, 10, dtype, stretched=True, anelastic=True)
    c0, c1 = copy.deepcopy(case), copy.deepcopy(case)
    N, R = both(g)
    prm = ostep.default_params(); prm.update(swadvec=swadvec, swdiff=swdiff, visc=0.5, svisc=0.7)
    ostep.dycore_step(g, N, c0, prm, 2.0)
    ostep.dycore_step(g, R, c1, prm, 2.0, pres=refbind.RefPres(g, 2, c1["rhoref"], c1["rhorefh"]))    # the reference's own Pres_2
    for n in ("u", "v", "w", "th", "p"):
        assert np.array_equal(c0[n], c1[n]), n
    assert np.isfinite(interior(g, c0["u"])).all()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,anel,stretched", CASES)
def test_full_rk3_step_bitexact(dtype, shape, anel, stretched):
    """One full RK3 step in the reference's call order: numpy oracle == compiled reference kernels."""
    g, gd, case = make_pair(*shape, dtype, stretched=stretched, anelastic=anel)
    c0, c1 = copy.deepcopy(case), copy.deepcopy(case)
    N, R = both(g)
    prm = ostep.default_params()
    ostep.dycore_step(g, N, c0, prm, 2.0)
    ostep.dycore_step(g, R, c1, prm, 2.0, pres=refbind.RefPres(g, 2, c1["rhoref"], c1["rhorefh"]))    # the reference's own Pres_2
    for n in ("u", "v", "w", "th", "p", "evisc"):
        assert np.array_equal(c0[n], c1[n]), n
    # the step does something and stays finite
    assert np.isfinite(interior(g, c0["u"])).all()
    assert not np.array_equal(c0["u"], case["u"])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("thermobc", [O.BC_FLUX, O.BC_DIRICHLET])
def test_boundary_surface_pinned(dtype, thermobc):
    """The Monin-Obukhov surface model (lookup solver, constant z0): the numpy restatement against the reference's own compiled
    kernels (stability, surfm, surfs, calc_dutot, calc_duvdz_mo, calc_dbdz_mo, prepare_lut) over two consecutive calls (the second
    starts the table search from the first one's index).  Transcendentals come from numpy vs libm: relative 1e-12 / 1e-5."""
    import copy
    from util import make_pair, prepare_halos, rel_l2
    g, gd, case = make_pair(24, 16, 8, dtype, stretched=True)
    prepare_halos(g, case)
    for n in ("u", "v", "th"):
        case[n + "_bot"] = np.zeros(gd.shape2d, dtype)
    case["th_bot"][...] = 300.5 if thermobc == O.BC_DIRICHLET else 0.
    a = copy.deepcopy(case); b = copy.deepcopy(case)
    S1 = O.BoundarySurface(g, 0.1, 0.01, O.BC_DIRICHLET, thermobc)
    S2 = refbind.RefSurface(g, 0.1, 0.01, O.BC_DIRICHLET, thermobc)
    assert np.array_equal(S1.zL_sl, S2.zL_sl)
    tol = 1e-12 if dtype == np.float64 else 1e-5
    assert rel_l2(S1.f_sl, S2.f_sl) <= (1e-7 if dtype == np.float64 else 1e-5)          # float table
    for it in range(2):
        d1 = S1.exec(a, case["thref"], case["threfh"]); d2 = S2.exec(b, case["thref"], case["threfh"])
    assert rel_l2(d1, d2) <= tol and rel_l2(S1.obuk, S2.obuk) <= tol and rel_l2(S1.ustar, S2.ustar) <= tol
    sl = (slice(g.jstart, g.jend), slice(g.istart, g.iend))
    for n in ("u_fluxbot", "v_fluxbot", "u_gradbot", "v_gradbot", "th_bot", "th_gradbot", "th_fluxbot"):
        assert rel_l2(a[n], b[n]) <= tol, n
    for n in ("dudz_mo", "dvdz_mo", "dbdz_mo"):
        assert rel_l2(a[n][sl], b[n][sl]) <= tol, n


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_buffer_and_force_pinned(dtype):
    """Damping layer and large-scale forcings: the numpy restatement against the reference's compiled kernels
    (calc_buffer, enforce_fixed_flux, calc_coriolis_2nd, calc_large_scale_source, advec_wls_2nd_local)."""
    from util import make_pair, prepare_halos, rel_l2
    g, gd, case = make_pair(24, 16, 12, dtype, stretched=True)
    prepare_halos(g, case)
    R = refbind.RefForcing(g)
    rng = np.random.default_rng(4)
    tol = 1e-13 if dtype == np.float64 else 1e-6
    prof = lambda: rng.standard_normal(gd.kcells).astype(dtype)
    # buffer (the damping factor uses pow: numpy vs libm)
    zstart = float(0.6*g.zsize)
    ks, ksh = O.buffer_kstart(g, zstart)
    assert g.kstart < ks < g.kend and ksh in (ks, ks + 1, ks - 1)
    for fld, z, k0 in (("u", g.z, ks), ("w", g.zh, ksh), ("th", g.z, ks)):
        abuf = prof()
        a = rng.standard_normal(gd.shape).astype(dtype); b = a.copy()
        O.calc_buffer(g, a, case[fld], abuf, z, zstart, 2., 2.5, k0)
        R.buffer(b, case[fld], abuf, z, zstart, 2., 2.5, k0)
        assert rel_l2(a, b) <= tol, fld
    # fixed mass flux
    ut = rng.standard_normal(gd.shape).astype(dtype); ut2 = ut.copy()
    um, utm = O.field_mean(g, case["u"]), O.field_mean(g, ut)
    O.force_fixed_flux(g, ut, case["u"], 0.11, 0.02, 0.7)
    R.fixed_flux(ut2, 0.11, um, utm, 0.02, 0.7)
    assert rel_l2(ut, ut2) <= tol
    # Coriolis + geostrophic wind
    ug, vg = prof(), prof()
    a = {n: rng.standard_normal(gd.shape).astype(dtype) for n in ("ut", "vt")}; b = copy.deepcopy(a)
    O.force_coriolis_2nd(g, a["ut"], a["vt"], case["u"], case["v"], ug, vg, 1e-4, 0.3, -0.2)
    R.coriolis(b["ut"], b["vt"], case["u"], case["v"], ug, vg, 1e-4, 0.3, -0.2)
    assert np.array_equal(a["ut"], b["ut"]) and np.array_equal(a["vt"], b["vt"])
    # large-scale source and local subsidence
    sls, wls = prof(), (0.01*prof())
    st = rng.standard_normal(gd.shape).astype(dtype); st2 = st.copy()
    O.force_ls_source(g, st, sls); O.force_wls_local(g, st, case["th"], wls)
    R.ls_source(st2, sls); R.wls_local(st2, case["th"], wls)
    assert np.array_equal(st, st2)


# ---------------------------------------------------------------------------------------------------------------------
# Tier-2 pin of the pressure glue: the reference's own FFT<TF>, Pres_2<TF>, Pres_4<TF> member functions (compiled where
# they lie, run on stand-in objects -- oracle/ref/ref_fake_pres.h) against the numpy restatement.  FFTW is absent from the
# image, so the plans the reference creates forward the batched 1-D R2HC / HC2R to the oracle's transform; everything
# around it -- slice loops and strides (src/fft.cxx:338-452), rhs, modified wave numbers, matrix build, tdma / hdma, ghost
# cells, pressure gradient -- is the reference's compiled code, and must agree bit for bit.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(16, 12, 8), (20, 9, 6), (24, 1, 8)])
def test_fft_glue_bitexact(dtype, shape):
    """FFT<TF>::init / load / exec_forward / exec_backward (src/fft.cxx): x then y transforms slice by slice."""
    g, gd, case = make_pair(*shape, dtype)
    R = refbind.RefPres(g, 2, case["rhoref"], case["rhorefh"])
    x = np.random.default_rng(0).standard_normal((g.kmax, g.jtot, g.itot)).astype(dtype)
    y = x.copy()
    R.fft_forward(x)
    y = O.r2hc(O.r2hc(y, 2), 1)
    assert np.array_equal(x, y)
    R.fft_backward(x)                     # normalises by jtot, itot (src/fft.cxx:424-449)
    y = O.hc2r(y, 1)/dtype(g.jtot)
    y = O.hc2r(y, 2)/dtype(g.itot)
    assert np.array_equal(x, y)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(16, 12, 8), (20, 9, 6), (24, 1, 8)])
@pytest.mark.parametrize("anelastic", [False, True])
def test_pres_2_glue_bitexact(dtype, shape, anelastic):
    """Pres_2::set_values / input / solve / output (src/pres_2.cxx:124-387) in the order of Pres_2::exec (:66-94)."""
    g, gd, case = make_pair(*shape, dtype, stretched=True, anelastic=anelastic)
    R = refbind.RefPres(g, 2, case["rhoref"], case["rhorefh"])
    P = O.Pres2(g, case["rhoref"], case["rhorefh"])
    for n in ("bmati", "bmatj", "a", "c"):
        assert np.array_equal(getattr(R, n), getattr(P, n)), n
    rng = np.random.default_rng(1)
    c0 = copy.deepcopy(case)
    for n in ("ut", "vt", "wt"):
        c0[n] = rng.standard_normal(gd.shape).astype(dtype)
    c1 = copy.deepcopy(c0)
    P.exec(c0["p"], c0["u"], c0["v"], c0["w"], c0["ut"], c0["vt"], c0["wt"], 0.7)
    R.exec(c1["p"], c1["u"], c1["v"], c1["w"], c1["ut"], c1["vt"], c1["wt"], 0.7)
    for n in ("ut", "vt", "wt"):
        assert np.array_equal(c0[n], c1[n]), (n, float(np.abs(c0[n].astype(np.float64) - c1[n]).max()))
    # p where solve() defines it: interior, the bottom ghost level and their cyclic ghost cells (elsewhere the reference's
    # array keeps scratch of the in-place compact rhs; nothing reads it)
    mask = R.defined_mask()
    assert np.array_equal(c0["p"][mask], c1["p"][mask])
    assert not np.array_equal(c0["ut"], case["ut"])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(16, 12, 10), (20, 9, 6), (24, 1, 8)])
@pytest.mark.parametrize("stretched", [False, True])
def test_pres_4_glue_bitexact(dtype, shape, stretched):
    """Pres_4::set_values / input<dim3> / solve (+ hdma) / output<dim3> / calc_divergence (src/pres_4.cxx:179-767) in the
    order and with the work-array carving of Pres_4::exec (:77-144), 3-D and 2-D (jtot = 1)."""
    from util import stretched_z
    it, jt, kt = shape
    g = O.Grid(it, jt, kt, 6.28, 3.14, 2., 3, 3, 3, dtype, z=stretched_z(kt, 2.) if stretched else None, order=4)
    R = refbind.RefPres(g, 4)
    P = O.Pres4(g)
    for n in ("bmati", "bmatj", "m"):
        assert np.array_equal(getattr(R, n), getattr(P, n)), n
    rng = np.random.default_rng(2)
    f = lambda: rng.standard_normal((g.kcells, g.jcells, g.icells)).astype(dtype)
    c0 = dict(p=f(), u=f(), v=f(), w=f(), ut=f(), vt=f(), wt=f())
    c1 = copy.deepcopy(c0)
    t0 = c0["ut"].copy()
    P.exec(c0["p"], c0["u"], c0["v"], c0["w"], c0["ut"], c0["vt"], c0["wt"], 0.7)
    R.exec(c1["p"], c1["u"], c1["v"], c1["w"], c1["ut"], c1["vt"], c1["wt"], 0.7)
    for n in ("ut", "vt", "wt"):
        assert np.array_equal(c0[n], c1[n]), (n, float(np.abs(c0[n].astype(np.float64) - c1[n]).max()))
    # p where solve() defines it: two ghost levels either side (read by output) and their cyclic ghost cells
    mask = R.defined_mask()
    assert np.array_equal(c0["p"][mask], c1["p"][mask])
    assert not np.array_equal(c0["ut"], t0)
    assert float(P.divergence(c0["u"], c0["v"], c0["w"])) == R.divergence(c1["u"], c1["v"], c1["w"])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("stretched", [False, True])
def test_grid_metrics_bitexact(dtype, order, stretched):
    """Grid<TF>::init + calculate (src/grid.cxx:106-376) compiled from the reference: ghost levels of z / zh, dz, dzh and their
    reciprocals, and the 4th-order metrics dzi4 / dzhi4 == the oracle's Grid AND the product's GridData, bit for bit."""
    from util import stretched_z
    from microhh_b200.grid import GridData
    it, jt, kt = 16, 12, 10
    gc = (3, 3, 3) if order == 4 else (3, 3, 1)
    z = stretched_z(kt, 2.) if stretched else None
    g = O.Grid(it, jt, kt, 6.28, 3.14, 2., *gc, dtype, z=z, order=order)
    gd = GridData(it, jt, kt, 6.28, 3.14, 2., *gc, dtype, z=z, order=order)
    if order == 4:
        R = refbind.RefPres(g, 4)
    else:
        ones = np.ones(g.kcells, dtype)
        R = refbind.RefPres(g, 2, ones, ones)
    m, scal = R.grid_metrics()
    names = ("z", "zh", "dz", "dzh", "dzi", "dzhi") + (("dzi4", "dzhi4") if order == 4 else ())
    for n in names:
        assert np.array_equal(np.asarray(getattr(g, n))[:g.kcells], m[n]), ("oracle", n)
        if hasattr(gd, n):
            assert np.array_equal(np.asarray(getattr(gd, n))[:g.kcells], m[n]), ("product", n)
    assert scal[0] == g.dx == gd.dx and scal[1] == g.dy == gd.dy


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,anel,stretched", CASES)
@pytest.mark.parametrize("mason", [True, False])
def test_diff_tke2_kernels_bitexact(dtype, shape, anel, stretched, mason):
    """Deardorff SGS-TKE closure (src/diff_tke2.cxx:48-512) and the limiter (src/limiter.cxx:35-59), kernel by kernel."""
    g, gd, case = make_pair(*shape, dtype, stretched=stretched, anelastic=anel)
    add_sgstke(g, case)
    prepare_halos(g, case)
    N, R = both(g)
    t = ostep.TKE2_DEFAULTS
    rng = np.random.default_rng(3)
    N2 = (1.e-4*rng.standard_normal(gd.shape)).astype(dtype)           # both signs: stable and unstable points
    str2 = (1.e-3*rng.random(gd.shape)).astype(dtype)
    seed_t = rng.standard_normal(gd.shape).astype(dtype)
    res = []
    for K in (N, R):
        c = copy.deepcopy(case)
        o = {}
        K.tke2_enforce_min(c["sgstke"]); o["sgstke"] = c["sgstke"].copy()
        ev = np.zeros(gd.shape, dtype); K.tke2_evisc_neutral(ev, c["sgstke"], c["u"], c["v"], c["w"], c["z0m"], t["cn"], t["cm"], mason); o["evisc_neutral"] = ev
        ev = np.zeros(gd.shape, dtype); K.tke2_evisc(ev, c["sgstke"], c["u"], c["v"], c["w"], N2, c["dbdz_mo"], c["z0m"], t["cn"], t["cm"], mason); o["evisc"] = ev
        evh = np.zeros(gd.shape, dtype); K.tke2_evisc_heat(evh, ev, c["sgstke"], N2, c["dbdz_mo"], c["z0m"], t["cn"], t["ch1"], t["ch2"], mason); o["eviscs"] = evh
        at = seed_t.copy(); K.tke2_shear_tend(at, c["sgstke"], ev, str2); o["shear"] = at
        at = seed_t.copy(); K.tke2_buoy_tend(at, c["sgstke"], evh, N2, c["dbdz_mo"]); o["buoy"] = at
        at = seed_t.copy(); K.tke2_diss_tend(at, c["sgstke"], N2, c["dbdz_mo"], c["z0m"], t["cn"], t["ce1"], t["ce2"], mason); o["diss"] = at
        at = seed_t.copy(); K.tke2_diss_tend_neutral(at, c["sgstke"], c["z0m"], t["ce1"], t["ce2"], mason); o["diss_neutral"] = at
        at = (-0.5*seed_t).copy(); K.tendency_limiter(at, c["sgstke"], O.SGSTKE_MIN, 1.7); o["limiter"] = at
        res.append(o)
    for n in res[0]:
        a, b = res[0][n], res[1][n]
        assert np.array_equal(a, b), (n, float(np.abs(a.astype(np.float64) - b).max()))
    assert np.isfinite(interior(g, res[0]["diss"])).all() and np.isfinite(interior(g, res[0]["evisc"])).all()
    assert (interior(g, res[0]["sgstke"]) >= dtype(O.SGSTKE_MIN)).all()
    assert not np.array_equal(res[0]["limiter"], -0.5*seed_t)           # the limiter engaged somewhere


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("swthermo", ["dry", None])
@pytest.mark.parametrize("swadvec", ["2i5", "2"])
def test_full_rk3_step_tke2_bitexact(dtype, swthermo, swadvec):
    """One RK3 step with swdiff = tke2 in Model::exec's order (exec_viscosity incl. the sgstke sources, diffusion with evisc /
    eviscs and tPr = 1, the limiter after the pressure solve): numpy oracle == compiled reference kernels."""
    g, gd, case = make_pair(20, 12, 10, dtype, stretched=True, anelastic=True, ns=2)
    add_sgstke(g, case)
    O.tke2_enforce_min(g, case["sgstke"])                               # Diff_tke2::create at a cold start
    c0, c1 = copy.deepcopy(case), copy.deepcopy(case)
    N, R = both(g)
    prm = ostep.default_params(); prm.update(swdiff="tke2", swthermo=swthermo, swadvec=swadvec)
    ostep.dycore_step(g, N, c0, prm, 2.0)
    ostep.dycore_step(g, R, c1, prm, 2.0, pres=refbind.RefPres(g, 2, c1["rhoref"], c1["rhorefh"]))    # the reference's own Pres_2
    for n in ("u", "v", "w", "th", "s1", "sgstke", "evisc", "eviscs", "p"):
        assert np.array_equal(c0[n], c1[n]), n
    assert np.isfinite(interior(g, c0["sgstke"])).all() and np.isfinite(interior(g, c0["u"])).all()
    assert not np.array_equal(c0["sgstke"], case["sgstke"])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,order", [((16, 12, 8), 2), ((24, 1, 8), 2), ((12, 10, 6), 4)])
@pytest.mark.parametrize("offset", [0., 300.])
def test_field3d_io_bitexact(dtype, shape, order, offset, tmp_path):
    """Restart IO (src/field3d_io.cxx:669-751): the oracle's file is byte for byte the reference's, either side loads the
    other's file, an existing file is refused (fopen "wbx") and a missing / short one reported."""
    g = O.Grid(*shape, 100., 80., 60., 3, 3, 3 if order == 4 else 1, dtype, order=order)
    rng = np.random.default_rng(9)
    a = rng.standard_normal(g.field().shape).astype(dtype)
    R = refbind.RefField3dIO(g)
    fo, fr = tmp_path / "u.oracle", tmp_path / "u.ref"
    assert O.field3d_save(g, a, str(fo), offset) == 0 and R.save(a.copy(), fr, offset) == 0
    assert fo.read_bytes() == fr.read_bytes()
    assert fo.stat().st_size == g.itot*g.jtot*g.ktot*np.dtype(dtype).itemsize
    assert O.field3d_save(g, a, str(fo), offset) != 0 and R.save(a.copy(), fr, offset) != 0          # exclusive create
    b0 = np.full_like(a, 7.); b1 = np.full_like(a, 7.)
    assert O.field3d_load(g, b0, str(fr), offset) == 0 and R.load(b1, fo, offset) == 0
    assert np.array_equal(b0, b1)
    assert np.array_equal(interior(g, b0), (interior(g, a) + dtype(offset)) - dtype(offset))
    assert (b0[0] == 7.).all()                                                                       # ghost cells untouched
    assert O.field3d_load(g, b0, str(tmp_path / "missing"), offset) != 0 and R.load(b1, tmp_path / "missing", offset) != 0


def _o4_pair(shape, dtype, stretched):
    """4th-order grid + synthetic case for both sides (as test_full_rk3_step_order4_bitexact)"""
    from microhh_b200.grid import GridData
    from microhh_b200.synthetic import make_case
    from util import stretched_z
    z = stretched_z(shape[2], 2.) if stretched else None
    g = O.Grid(*shape, 2*np.pi, np.pi, 2., 3, 3, 3, dtype, z=z, order=4)
    gd = GridData(*shape, 2*np.pi, np.pi, 2., 3, 3, 3, dtype, z=z, order=4)
    case = make_case(gd, seed=5, noise=0.02)
    case["w"][:g.kstart+1] = 0; case["w"][g.kend:] = 0
    for n in ("u", "v"):
        for sfx in ("_bot", "_top", "_gradbot", "_gradtop"):
            case[n + sfx] = np.zeros(gd.shape2d, dtype)
    case["th_gradbot"] = np.full(gd.shape2d, -0.3, dtype); case["th_gradtop"] = np.full(gd.shape2d, 0.2, dtype)
    return g, gd, case


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("stretched", [False, True])
def test_thermo_buoy_kernels_bitexact(dtype, order, stretched):
    """Thermo_buoy (src/thermo_buoy.cxx:41-296): N2, buoyancy tendency, slope variants, baroclinic term, 2nd and 4th order."""
    if order == 4:
        g, gd, case = _o4_pair((16, 12, 10), dtype, stretched)
    else:
        g, gd, case = make_pair(16, 12, 10, dtype, stretched=stretched)
    N, R = both(g)
    rng = np.random.default_rng(17)
    b = (0.1*rng.standard_normal(gd.shape)).astype(dtype)
    seeds = {n: rng.standard_normal(gd.shape).astype(dtype) for n in ("ut", "wt", "bt")}
    res = []
    for K in (N, R):
        o = {}
        n2 = np.zeros(gd.shape, dtype); K.thermo_buoy_N2(n2, b, 1.5e-4); o["N2"] = n2
        wt = seeds["wt"].copy(); K.thermo_buoy_tend(wt, b, order); o["wt"] = wt
        ut, wt, bt = seeds["ut"].copy(), seeds["wt"].copy(), seeds["bt"].copy()
        K.thermo_buoy_tend_slope(ut, wt, bt, b, case["u"], case["w"], 0.17, 2.e-3, 0.4, order)
        o["ut_s"], o["wt_s"], o["bt_s"] = ut, wt, bt
        bt = seeds["bt"].copy(); K.thermo_buoy_baroclinic(bt, case["v"], 3.e-4, order); o["bt_b"] = bt
        res.append(o)
    for n in res[0]:
        assert np.array_equal(res[0][n], res[1][n]), (n, float(np.abs(res[0][n].astype(np.float64) - res[1][n]).max()))
    assert not np.array_equal(res[0]["bt_s"], seeds["bt"])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("tb", [{}, dict(alpha=0.2, n2=1.e-2, utrans=0.1), dict(swbaroclinic=True, dbdy_ls=2.e-3)])
def test_full_rk3_step_order4_buoy_bitexact(dtype, tb):
    """The 4th-order DNS step with swthermo = buoy (cases/drycbl, rayleighbenard, weakscaling ...): thermo.exec between the
    ghost cells and the advection, plain / slope-enabled / baroclinic."""
    g, gd, case = _o4_pair((16, 12, 10), dtype, True)
    case["scalars"] = ["th"]            # scalar 0 is the buoyancy
    case["th"] = (case["th"] - dtype(300.)).astype(dtype)
    c0, c1 = copy.deepcopy(case), copy.deepcopy(case)
    N, R = both(g)
    prm = ostep.default_params(); prm.update(swadvec="4m", swdiff="4", swthermo="buoy", thermo_buoy=tb, visc=1e-3, svisc=1e-3,
                                             mbcbot=0, mbctop=0)
    ostep.dycore_step(g, N, c0, prm, 1e-3)
    ostep.dycore_step(g, R, c1, prm, 1e-3, pres=refbind.RefPres(g, 4))
    for n in ("u", "v", "w", "th", "p"):
        assert np.array_equal(c0[n], c1[n]), n
    c2 = copy.deepcopy(case)
    prm2 = dict(prm); prm2.update(swthermo=None)
    ostep.dycore_step(g, N, c2, prm2, 1e-3)
    assert not np.array_equal(c0["w"], c2["w"])                          # the buoyancy does act


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("tb", [{}, dict(alpha=0.2, n2=1.e-2, utrans=0.1, swbaroclinic=True, dbdy_ls=2.e-3)])
def test_full_rk3_step_order2_buoy_bitexact(dtype, tb):
    """swthermo = buoy on a 2nd-order grid (Thermo_buoy::exec's Grid_order::Second branch) with advec_2i5 + diff_2."""
    g, gd, case = make_pair(16, 12, 10, dtype, stretched=True)
    case["th"] = (case["th"] - dtype(300.)).astype(dtype)
    c0, c1 = copy.deepcopy(case), copy.deepcopy(case)
    N, R = both(g)
    prm = ostep.default_params(); prm.update(swdiff="2", swthermo="buoy", thermo_buoy=tb, surface_model=False, visc=1e-2, svisc=1e-2)
    ostep.dycore_step(g, N, c0, prm, 1e-2)
    ostep.dycore_step(g, R, c1, prm, 1e-2, pres=refbind.RefPres(g, 2, c1["rhoref"], c1["rhorefh"]))
    for n in ("u", "v", "w", "th", "p"):
        assert np.array_equal(c0[n], c1[n]), n
    c2 = copy.deepcopy(case)
    prm2 = dict(prm); prm2.update(swthermo=None)
    ostep.dycore_step(g, N, c2, prm2, 1e-2)
    assert not np.array_equal(c0["w"], c2["w"])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cold", [False, True])
def test_thermo_moist_bitexact(dtype, cold):
    """Thermo_moist (src/thermo_moist.cxx, include/thermo_moist_functions.h): base state, saturation adjustment inside the
    buoyancy tendency, the b / ql / N2 diagnostics and the surface buoyancy helpers, warm and mixed-phase."""
    g, gd, case = make_pair(16, 12, 24, dtype, stretched=True, sizes=(3200., 3200., 3000.))
    N, R = both(g)
    thl, qt = moist_case(g, gd, dtype, cold=cold)
    rng = np.random.default_rng(23)
    # mean profiles (ghost levels included) -> reference profiles with calc_top_and_bot ghosts -> base state
    mt = [K.mean_profile(thl) for K in (N, R)]; mq = [K.mean_profile(qt) for K in (N, R)]
    assert np.array_equal(mt[0], mt[1]) and np.array_equal(mq[0], mq[1])
    tb = []
    for K in (N, R):
        a, b = mt[0].copy(), mq[0].copy(); K.moist_top_and_bot(a, b); tb.append((a, b))
    assert np.array_equal(tb[0][0], tb[1][0]) and np.array_equal(tb[0][1], tb[1][1])
    pbot = 101500. if not cold else 70000.
    bs = [K.moist_base_state(tb[0][0], tb[0][1], pbot) for K in (N, R)]
    for n in bs[0]:
        assert np.array_equal(bs[0][n], bs[1][n]), (n, bs[0][n], bs[1][n])
    assert 0.9 < bs[0]["rhorefh"][g.kstart] < 1.3 or cold
    b0 = bs[0]
    res = []
    for K in (N, R):
        o = {}
        wt = rng.standard_normal(gd.shape).astype(dtype) if not res else res[0]["wt0"].copy()
        o["wt0"] = wt.copy()
        K.thermo_moist_buoyancy_tend_2nd(wt, thl, qt, b0["prefh"], b0["thvrefh"]); o["wt"] = wt
        b = np.zeros(gd.shape, dtype); K.thermo_moist_buoyancy(b, thl, qt, b0["pref"], b0["thvref"] + (b0["thvref"] == 0)*dtype(300.)); o["b"] = b
        ql = np.zeros(gd.shape, dtype); K.thermo_moist_liquid_water(ql, thl, qt, b0["pref"]); o["ql"] = ql
        n2 = np.zeros(gd.shape, dtype); K.thermo_moist_N2(n2, thl, b0["thvref"]); o["N2"] = n2
        bb = np.zeros(gd.shape, dtype); bbot = np.zeros(gd.shape2d, dtype)
        K.thermo_moist_buoyancy_bot(bb, bbot, thl, thl[g.kstart] + dtype(0.5), qt, qt[g.kstart]*dtype(1.1), b0["thvref"], b0["thvrefh"])
        o["b_ks"], o["bbot"] = bb, bbot
        bf = np.zeros(gd.shape2d, dtype)
        K.thermo_moist_buoyancy_fluxbot(bf, thl, np.full(gd.shape2d, 8.e-3, dtype), qt, np.full(gd.shape2d, 5.2e-5, dtype), b0["thvrefh"])
        o["bfluxbot"] = bf
        res.append(o)
    for n in res[0]:
        assert np.array_equal(res[0][n], res[1][n]), (n, float(np.abs(res[0][n].astype(np.float64) - res[1][n]).max()))
    frac = float((interior(g, res[0]["ql"]) > 0).mean())
    assert 0.05 < frac < 0.95, frac                       # both branches of the adjustment are exercised


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("swdiff,update", [("smag2", True), ("smag2", False), ("2", True)])
def test_full_rk3_step_moist_bitexact(dtype, swdiff, update):
    """One RK3 step with swthermo = moist (bomex-like state): N2 of the closure from thvref, the base state re-integrated from
    the mean profiles in every sub-step (swupdatebasestate), buoyancy through the saturation adjustment."""
    g, gd, case, pbot = make_moist_pair(16, 12, 20, dtype)
    c0, c1 = copy.deepcopy(case), copy.deepcopy(case)
    N, R = both(g)
    prm = ostep.default_params(); prm.update(swdiff=swdiff, swthermo="moist", thermo_moist=dict(pbot=pbot, swupdatebasestate=update),
                                             surface_model=(swdiff == "smag2"), visc=1e-5 if swdiff == "smag2" else 1e-2, svisc=1e-5 if swdiff == "smag2" else 1e-2)
    ostep.dycore_step(g, N, c0, prm, 2.0)
    ostep.dycore_step(g, R, c1, prm, 2.0, pres=refbind.RefPres(g, 2, c1["rhoref"], c1["rhorefh"]))
    for n in ("u", "v", "w", "thl", "qt", "p"):
        assert np.array_equal(c0[n], c1[n]), n
    for n in c0["moist_bs"]:
        assert np.array_equal(c0["moist_bs"][n], c1["moist_bs"][n]), n
    assert update == (not np.array_equal(c0["moist_bs"]["thvrefh"], case["moist_bs"]["thvrefh"]))
    c2 = copy.deepcopy(case)
    prm2 = dict(prm); prm2.update(swthermo=None)
    ostep.dycore_step(g, N, c2, prm2, 2.0)
    assert not np.array_equal(c0["w"], c2["w"])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("scheme,gc", [("2i4", (2, 2, 2)), ("2i62", (3, 3, 1))])
@pytest.mark.parametrize("shape", [(16, 12, 10), (12, 1, 8)])
@pytest.mark.parametrize("anel", [False, True])
def test_advec_2i4_2i62_bitexact(dtype, scheme, gc, shape, anel):
    """Advec_2i4 (src/advec_2i4.cxx:53-518) and Advec_2i62 (src/advec_2i62.cxx:59-306): u, v, w, scalar tendencies and the CFL
    number, with the ghost cells each scheme asks for, Boussinesq and anelastic, stretched grid."""
    from microhh_b200.grid import GridData
    from microhh_b200.synthetic import make_case
    from util import stretched_z
    z = stretched_z(shape[2], 3200.)
    g = O.Grid(*shape, 3200., 3200., 3200., *gc, dtype, z=z)
    gd = GridData(*shape, 3200., 3200., 3200., *gc, dtype, z=z)
    case = make_case(gd, seed=4, anelastic=anel)
    N, R = both(g)
    rng = np.random.default_rng(9)
    for n in ("u", "v", "w", "th"):
        case[n] = (case[n] + 0.1*rng.standard_normal(gd.shape)).astype(dtype)         # ghost cells too: both sides read the same arrays
    res = []
    for K in (N, R):
        o = {n: np.zeros(gd.shape, dtype) for n in ("ut", "vt", "wt", "tht")}
        a = (case["u"], case["v"], case["w"], case["rhoref"], case["rhorefh"])
        getattr(K, f"advec_{scheme}_u")(o["ut"], *a); getattr(K, f"advec_{scheme}_v")(o["vt"], *a); getattr(K, f"advec_{scheme}_w")(o["wt"], *a)
        getattr(K, f"advec_{scheme}_s")(o["tht"], case["th"], *a)
        o["cfl"] = np.array(getattr(K, f"advec_{scheme}_cfl")(case["u"], case["v"], case["w"], 3.0))
        res.append(o)
    for n in res[0]:
        assert np.array_equal(res[0][n], res[1][n]), (n, float(np.abs(res[0][n].astype(np.float64) - res[1][n]).max()))
    assert np.abs(res[0]["ut"]).max() > 0 and res[0]["cfl"] > 0
    assert (res[0]["wt"][g.kstart] == 0).all() and (res[0]["wt"][g.kend:] == 0).all()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("scheme,gc,swdiff", [("2i4", (2, 2, 2), "smag2"), ("2i62", (3, 3, 1), "smag2"), ("2i62", (3, 3, 2), "2")])
def test_full_rk3_step_2i4_2i62_bitexact(dtype, scheme, gc, swdiff):
    """Full RK3 steps with Advec_2i4 (cases/gabls4s3: + smag2 + dry) and Advec_2i62 (+ a flux-limited second scalar)."""
    from microhh_b200.grid import GridData
    from microhh_b200.synthetic import make_case
    from util import stretched_z
    shape = (16, 12, 10)
    z = stretched_z(shape[2], 3200.)
    g = O.Grid(*shape, 3200., 3200., 3200., *gc, dtype, z=z)
    gd = GridData(*shape, 3200., 3200., 3200., *gc, dtype, z=z)
    case = make_case(gd, seed=4, anelastic=True, ns=2)
    c0, c1 = copy.deepcopy(case), copy.deepcopy(case)
    N, R = both(g)
    smag = swdiff == "smag2"
    prm = ostep.default_params(); prm.update(swadvec=scheme, swdiff=swdiff, surface_model=smag, visc=1e-5 if smag else 1e-2, svisc=1e-5 if smag else 1e-2,
                                             fluxlimit_list=("s1",) if gc[2] == 2 and scheme == "2i62" else ())
    ostep.dycore_step(g, N, c0, prm, 2.0)
    ostep.dycore_step(g, R, c1, prm, 2.0, pres=refbind.RefPres(g, 2, c1["rhoref"], c1["rhorefh"]))
    for n in ("u", "v", "w", "th", "s1", "p"):
        assert np.array_equal(c0[n], c1[n]), n
