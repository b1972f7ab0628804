"""
cuFFT as the correctness comparator of the library's own transforms (BASELINE.json north_star: "cuFFT serves only as a
correctness and performance comparator").  `torch.fft.rfft2 / irfft2` are cuFFT (D2Z / Z2D plans); the per-mode Thomas
solve between them is a plain torch restatement of the reference's tdma (src/pres_2.cxx:202-324).  Unlike the numpy oracle
this comparator runs at the lengths bench.py times: 512 x 512 (N = 1) and the 1024-point transforms of the 8-GPU run.
The timing side of the comparison is tools/fft_compare.py.
"""
import numpy as np
import pytest

from util import TOL, rel_l2, make_pair
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def cufft_solve(torch, rhs, P, g, rhoref):
    """rfft2 (cuFFT) -> tridiagonal solve per (l, m) mode -> irfft2 (cuFFT); rhs: (k, j, i) device tensor."""
    tdt = rhs.dtype
    dev = rhs.device
    nm = g.itot//2 + 1
    spec = torch.fft.rfft2(rhs, dim=(1, 2))                                   # (k, l, m), m = 0..itot/2
    bmati = torch.from_numpy(P.bmati[:nm].copy()).to(dev)
    bmatj = torch.from_numpy(P.bmatj.copy()).to(dev)
    a = torch.from_numpy(P.a.copy()).to(dev); c = torch.from_numpy(P.c.copy()).to(dev)
    kgc = g.kgc
    dz = torch.from_numpy(g.dz[kgc:kgc+g.kmax].copy()).to(dev)
    rho = torch.from_numpy(rhoref[kgc:kgc+g.kmax].copy()).to(dev)
    lam = bmati[None, :] + bmatj[:, None]                                     # (l, m)
    K = g.kmax
    b = [dz[k]*dz[k]*rho[k]*lam - (a[k] + c[k]) for k in range(K)]
    b[0] = b[0] + a[0]
    top = torch.full_like(lam, float(P.c[K-1])); top[0, 0] = -float(P.c[K-1])
    b[K-1] = b[K-1] + top
    p = [dz[k]*dz[k]*spec[k] for k in range(K)]
    work3d = [None]*K
    w2 = b[0]
    p[0] = p[0]/w2
    for k in range(1, K):
        work3d[k] = c[k-1]/w2
        w2 = b[k] - a[k]*work3d[k]
        p[k] = (p[k] - a[k]*p[k-1])/w2
    for k in range(K-2, -1, -1):
        p[k] = p[k] - work3d[k+1]*p[k+1]
    out = torch.fft.irfft2(torch.stack(p), s=(g.jtot, g.itot), dim=(1, 2))
    return out.to(tdt)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(512, 512, 24), (1024, 1024, 8), (256, 1024, 8), (1024, 64, 8), (2048, 32, 6), (64, 2048, 6)])
def test_solve_matches_cufft_at_bench_lengths(dtype, shape):
    import torch
    from microhh_b200 import dycore as D
    g, gd, case = make_pair(*shape, dtype, stretched=True, anelastic=True)
    ctx = D.Context(gd, 0)
    ctx.set_basestate(case["rhoref"], case["rhorefh"], case["thref"], case["threfh"])
    rng = np.random.default_rng(shape[0] + shape[1])
    rhs = rng.standard_normal((gd.kmax, gd.jmax, gd.imax)).astype(dtype)
    a_in = torch.from_numpy(rhs).cuda(); a_out = torch.zeros_like(a_in)
    pres = D.Pres(ctx)
    # transforms alone: forward x, forward y, inverse y, inverse x == identity
    pres.fft_roundtrip(a_in, a_out, solve=False)
    assert rel_l2(a_out.cpu().numpy(), rhs) <= 20*TOL[dtype]
    # with the solve: against cuFFT + torch Thomas
    pres.fft_roundtrip(a_in, a_out, solve=True)
    P = O.Pres2(g, case["rhoref"], case["rhorefh"])
    ref = cufft_solve(torch, a_in, P, g, case["rhoref"])
    # fp32: two independent fp32 solves of an ill-conditioned system (the mean mode) agree to ~1e-4 at these sizes
    tol = 50*TOL[dtype] if dtype == np.float64 else 2e-3
    assert rel_l2(a_out.cpu().numpy(), ref.cpu().numpy()) <= tol


@pytest.mark.parametrize("n", [512, 1024])
def test_forward_spectrum_matches_cufft(n):
    """The forward x transform alone (real-to-half-complex of every row) against cuFFT, through the spectral round trip of a
    field that is a single y-mode: the library's x spectrum must then be cuFFT's spectrum of that row."""
    import torch
    from microhh_b200 import dycore as D
    dtype = np.float64
    g, gd, case = make_pair(n, 8, 6, dtype)
    ctx = D.Context(gd, 0)
    ctx.set_basestate(case["rhoref"], case["rhorefh"], case["thref"], case["threfh"])
    rng = np.random.default_rng(n)
    row = rng.standard_normal(n)
    rhs = np.broadcast_to(row, (gd.kmax, gd.jmax, gd.imax)).copy()
    a_in = torch.from_numpy(rhs).cuda(); a_out = torch.zeros_like(a_in)
    D.Pres(ctx).fft_roundtrip(a_in, a_out, solve=False)
    spec = torch.fft.rfft(a_in[0, 0])
    back = torch.fft.irfft(spec, n=n)
    assert rel_l2(a_out[0, 0].cpu().numpy(), back.cpu().numpy()) <= 1e-13
