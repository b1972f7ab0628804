"""
Golden vectors (tests/golden/*.npz, made by tests/make_golden.py from the REFERENCE's own compiled
CPU kernels).  CPU part: the numpy oracle reproduces them bit for bit and the FFT restatement obeys
FFTW's documented R2HC/HC2R definition.  GPU part: the CUDA path through the C ABI matches them
within BASELINE.json's tolerance (rel. L2 <= 1e-12 fp64 / 1e-5 fp32).
"""
import copy
import glob
import os

import numpy as np
import pytest

from util import TOL, rel_l2, make_pair, prepare_halos, interior
from make_golden import VISC2, SVISC2, input_digest, make_case4, params4, input_digest4, VISC4, DT
from oracle import oracle as O, step as ostep

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_*.npz")))
GOLD4 = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step4_*.npz")))


def load(path):
    z = np.load(path)
    dtype = np.float32 if path.endswith("_f32.npz") else np.float64
    g, gd, case = make_pair(*[int(x) for x in z["shape"]], dtype, stretched=bool(z["stretched"]), anelastic=bool(z["anel"]))
    assert input_digest(case) == str(z["input_sha256"]), "synthetic input generator drifted from the golden inputs"
    return z, dtype, g, gd, case


def test_golden_present():
    assert len(GOLD) >= 5


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_reproduces_golden_bitexact(path):
    z, dtype, g, gd, case = load(path)
    K = O.NumpyKernels(g)
    ck = copy.deepcopy(case); prepare_halos(g, ck)
    rr, rh = ck["rhoref"], ck["rhorefh"]
    K.advec_2i5_u(ck["ut"], ck["u"], ck["v"], ck["w"], rr, rh)
    K.advec_2i5_v(ck["vt"], ck["u"], ck["v"], ck["w"], rr, rh)
    K.advec_2i5_w(ck["wt"], ck["u"], ck["v"], ck["w"], rr, rh)
    K.advec_2i5_s(ck["tht"], ck["th"], ck["u"], ck["v"], ck["w"], rr, rh)
    for n in ("ut", "vt", "wt", "tht"):
        assert np.array_equal(ck[n], z["advec_" + n]), n
    assert K.advec_2i5_cfl(ck["u"], ck["v"], ck["w"], float(z["dt"])) == float(z["cfl"])
    K.diff_strain2(ck["evisc"], ck["u"], ck["v"], ck["w"], ck["dudz_mo"], ck["dvdz_mo"], True)
    n2 = np.zeros_like(ck["evisc"]); K.thermo_dry_N2(n2, ck["th"], ck["thref"])
    K.diff_evisc(ck["evisc"], ck["u"], ck["v"], ck["w"], n2, ck["dbdz_mo"], ck["z0m"], 0.23, 1./3., True, True)
    assert np.array_equal(ck["evisc"], z["evisc"])
    cs = copy.deepcopy(case)
    for _ in range(int(z["nsteps"])):
        ostep.dycore_step(g, K, cs, ostep.default_params(), float(z["dt"]))
    for n in ("u", "v", "w", "th", "p"):
        assert np.array_equal(cs[n], z["step_" + n]), n
    # Advec_2 / Diff_2 vectors
    c2 = copy.deepcopy(case); prepare_halos(g, c2)
    K.advec_2_u(c2["ut"], c2["u"], c2["v"], c2["w"], rr, rh)
    K.advec_2_v(c2["vt"], c2["u"], c2["v"], c2["w"], rr, rh)
    K.advec_2_w(c2["wt"], c2["u"], c2["v"], c2["w"], rr, rh)
    K.advec_2_s(c2["tht"], c2["th"], c2["u"], c2["v"], c2["w"], rr, rh)
    for n in ("ut", "vt", "wt", "tht"):
        assert np.array_equal(c2[n], z["advec2_" + n]), n
    assert K.advec_2_cfl(c2["u"], c2["v"], c2["w"], float(z["dt"])) == float(z["cfl2"])
    c3 = copy.deepcopy(case); prepare_halos(g, c3)
    K.diff_2_c(c3["ut"], c3["u"], VISC2); K.diff_2_c(c3["vt"], c3["v"], VISC2); K.diff_2_w(c3["wt"], c3["w"], VISC2)
    K.diff_2_c(c3["tht"], c3["th"], SVISC2)
    for n in ("ut", "vt", "wt", "tht"):
        assert np.array_equal(c3[n], z["diff2_" + n]), n
    c4 = copy.deepcopy(case)
    prm2 = ostep.default_params(); prm2.update(swadvec="2", swdiff="2", visc=VISC2, svisc=SVISC2)
    ostep.dycore_step(g, K, c4, prm2, float(z["dt"]))
    for n in ("u", "v", "w", "th"):
        assert np.array_equal(c4[n], z["step22_" + n]), n


@pytest.mark.parametrize("n", [2, 6, 8, 12, 20, 30])
def test_fft_restatement_obeys_fftw_definition(n):
    """FFTW r2r kinds used by the reference (src/fft.cxx:145-155): R2HC stores
    [r0, r1, ..., r_{n/2}, i_{(n+1)/2-1}, ..., i_1] of X_k = sum_j x_j exp(-2 pi i j k / n); HC2R is its
    unnormalised inverse (fftw3 manual, 'The Halfcomplex-format DFT')."""
    rng = np.random.default_rng(n)
    x = rng.standard_normal((3, n))
    j = np.arange(n)
    X = np.array([[np.sum(row*np.exp(-2j*np.pi*j*k/n)) for k in range(n)] for row in x])
    hc = O.r2hc(x.copy(), axis=1)
    exp = np.zeros_like(x)
    exp[:, :n//2 + 1] = X[:, :n//2 + 1].real
    for k in range(1, (n + 1)//2):
        exp[:, n - k] = X[:, k].imag
    assert np.allclose(hc, exp, rtol=0, atol=1e-12*n)
    back = O.hc2r(hc.copy(), axis=1)
    assert np.allclose(back, n*x, rtol=0, atol=1e-12*n*n)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_cuda_matches_golden(path):
    import torch
    from microhh_b200 import dycore as D
    z, dtype, g, gd, case = load(path)
    ctx = D.Context(gd, 0)
    ctx.set_basestate(case["rhoref"], case["rhorefh"], case["thref"], case["threfh"])
    # single kernels on halo-filled inputs
    ck = copy.deepcopy(case); prepare_halos(g, ck)
    f = D.Fields(ctx, ck)
    prm = D.make_params()
    D.Advec(ctx).exec(f)
    for n in ("ut", "vt", "wt", "tht"):
        k0 = g.kstart + 1 if n == "wt" else g.kstart
        assert rel_l2(interior(g, f[n].cpu().numpy(), k0), interior(g, z["advec_" + n], k0)) <= TOL[dtype], n
    cfl = D.Advec(ctx).get_cfl(f, float(z["dt"]))
    assert abs(cfl - float(z["cfl"])) <= 10*TOL[dtype]*float(z["cfl"])
    D.Diff(ctx, prm).exec_viscosity(f)
    assert rel_l2(f["evisc"].cpu().numpy(), z["evisc"]) <= 10*TOL[dtype]
    dn = D.Diff(ctx, prm).get_dn(f, float(z["dt"]))
    assert abs(dn - float(z["dn"])) <= 100*TOL[dtype]*float(z["dn"])
    # Advec_2 / Diff_2 kernels and the (2, 2) full step
    f3 = D.Fields(ctx, ck)
    D.Advec(ctx, "2").exec(f3)
    for n in ("ut", "vt", "wt", "tht"):
        k0 = g.kstart + 1 if n == "wt" else g.kstart
        assert rel_l2(interior(g, f3[n].cpu().numpy(), k0), interior(g, z["advec2_" + n], k0)) <= TOL[dtype], n
    cfl2 = D.Advec(ctx, "2").get_cfl(f3, float(z["dt"]))
    assert abs(cfl2 - float(z["cfl2"])) <= 10*TOL[dtype]*float(z["cfl2"])
    f4 = D.Fields(ctx, ck, visc=VISC2, svisc=SVISC2)
    D.Diff_2(ctx).exec(f4)
    for n in ("ut", "vt", "wt", "tht"):
        k0 = g.kstart + 1 if n == "wt" else g.kstart
        assert rel_l2(interior(g, f4[n].cpu().numpy(), k0), interior(g, z["diff2_" + n], k0)) <= TOL[dtype], n
    f5 = D.Fields(ctx, case, visc=VISC2, svisc=SVISC2)
    D.Dycore(ctx, D.make_params(swadvec="2", swdiff="2")).step(f5, float(z["dt"]))
    ctx.sync()
    for n in ("u", "v", "w", "th"):
        assert rel_l2(interior(g, f5[n].cpu().numpy()), interior(g, z["step22_" + n])) <= TOL[dtype], n
    # full step
    f2 = D.Fields(ctx, case)
    for _ in range(int(z["nsteps"])):
        D.Dycore(ctx, prm).step(f2, float(z["dt"]))
    ctx.sync()
    for n in ("u", "v", "w", "th"):
        assert rel_l2(interior(g, f2[n].cpu().numpy()), interior(g, z["step_" + n])) <= TOL[dtype], n


# ---- 4th-order DNS configuration (advec_4 + diff_4 + pres_4 + 4th-order ghost cells) -----------------------------
def load4(path):
    z = np.load(path)
    dtype = np.float32 if path.endswith("_f32.npz") else np.float64
    g, gd, case = make_case4(tuple(int(x) for x in z["shape"]), dtype)
    assert input_digest4(case) == str(z["input_sha256"]), "synthetic input generator drifted from the golden inputs"
    return z, dtype, g, gd, case


def test_golden4_present():
    assert len(GOLD4) >= 3


@pytest.mark.parametrize("path", GOLD4, ids=[os.path.basename(p)[:-4] for p in GOLD4])
def test_oracle_reproduces_golden4_bitexact(path):
    z, dtype, g, gd, case = load4(path)
    K = O.NumpyKernels(g)
    ck = copy.deepcopy(case)
    for n in ("u", "v", "w", "th"):
        K.boundary_cyclic(ck[n])
    K.advec_4_u(ck["ut"], ck["u"], ck["v"], ck["w"]); K.advec_4_v(ck["vt"], ck["u"], ck["v"], ck["w"])
    K.advec_4_w(ck["wt"], ck["u"], ck["v"], ck["w"]); K.advec_4_s(ck["tht"], ck["th"], ck["u"], ck["v"], ck["w"])
    for n in ("ut", "vt", "wt", "tht"):
        assert np.array_equal(ck[n], z["advec4_" + n]), n
    assert K.advec_4_cfl(ck["u"], ck["v"], ck["w"], DT) == float(z["cfl4"])
    K.diff_4_c(ck["ut"], ck["u"], VISC4); K.diff_4_c(ck["vt"], ck["v"], VISC4); K.diff_4_w(ck["wt"], ck["w"], VISC4)
    K.diff_4_c(ck["tht"], ck["th"], VISC4)
    for n in ("ut", "vt", "wt", "tht"):
        assert np.array_equal(ck[n], z["advdiff4_" + n]), n
    cs = copy.deepcopy(case)
    ostep.dycore_step(g, K, cs, params4(int(z["mbc"])), float(z["dt"]))
    for n in ("u", "v", "w", "th", "p"):
        assert np.array_equal(interior(g, cs[n]), interior(g, z["step_" + n])), n


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD4, ids=[os.path.basename(p)[:-4] for p in GOLD4])
def test_cuda_matches_golden4(path):
    from microhh_b200 import dycore as D
    z, dtype, g, gd, case = load4(path)
    ctx = D.Context(gd, 0)
    ones = np.ones(gd.kcells, dtype)
    ctx.set_basestate(ones, ones, 300*ones, 300*ones)
    ck = copy.deepcopy(case)
    for n in ("u", "v", "w", "th"):
        O.boundary_cyclic(g, ck[n])
    f = D.Fields(ctx, ck, visc=VISC4, svisc=VISC4)
    D.Advec(ctx, "4").exec(f)
    two_d = g.jtot == 1
    for n in ("ut", "vt", "wt", "tht"):
        if two_d and n == "vt":
            continue
        k0 = g.kstart + 1 if n == "wt" else g.kstart
        assert rel_l2(interior(g, f[n].cpu().numpy(), k0), interior(g, z["advec4_" + n], k0)) <= TOL[dtype], n
    cfl = D.Advec(ctx, "4").get_cfl(f, DT)
    assert abs(cfl - float(z["cfl4"])) <= 10*TOL[dtype]*float(z["cfl4"])
    D.Diff_4(ctx).exec(f)
    for n in ("ut", "vt", "wt", "tht"):
        if two_d and n == "vt":
            continue
        k0 = g.kstart + 1 if n == "wt" else g.kstart
        assert rel_l2(interior(g, f[n].cpu().numpy(), k0), interior(g, z["advdiff4_" + n], k0)) <= TOL[dtype], n
    mbc = int(z["mbc"])
    f2 = D.Fields(ctx, case, visc=VISC4, svisc=VISC4)
    prm = D.make_params(swadvec="4", swdiff="4", swthermo=None, surface_model=False, mbcbot=mbc, mbctop=mbc)
    D.Dycore(ctx, prm).step(f2, float(z["dt"]))
    ctx.sync()
    for n in (("u", "w", "th") if two_d else ("u", "v", "w", "th")):
        assert rel_l2(interior(g, f2[n].cpu().numpy()), interior(g, z["step_" + n])) <= 20*TOL[dtype], n
