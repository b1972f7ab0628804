"""Pres_4 oracle (oracle.Pres4, restating src/pres_4.cxx:178-767).  The reference's Pres_4 consists of class member
functions that need live Grid/Fields/Master objects, so there is no compiled-reference pin; the restatement is pinned by
the scheme's defining properties instead: the corrected velocities are divergence-free under the 4th-order divergence to
rounding, the solve inverts the discrete 4th-order div(grad) operator, and the banded LU (`hdma`) solves its system."""
import numpy as np
import pytest

from oracle import oracle as O
from util import stretched_z

EPS = {np.float64: 1e-12, np.float32: 5e-4}


def wall_bounded_velocity(g, rng, dtype):
    fld = lambda: rng.standard_normal((g.kcells, g.jcells, g.icells)).astype(dtype)
    u, v, w = fld(), fld(), fld()
    if g.jtot == 1:
        v[:] = 0
    ks, ke = g.kstart, g.kend
    w[ks] = 0; w[ke] = 0; w[ks-1] = -w[ks+1]; w[ke+1] = -w[ke-1]
    for a in (u, v, w):
        O.boundary_cyclic(g, a)
    return u, v, w


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(16, 12, 10), (24, 1, 8), (20, 10, 6)])
@pytest.mark.parametrize("stretched", [False, True])
def test_pres4_projects_to_divergence_free(dtype, shape, stretched):
    it, jt, kt = shape
    g = O.Grid(it, jt, kt, 6., 4., 2., 3, 3, 3, dtype, z=stretched_z(kt, 2.) if stretched else None, order=4)
    rng = np.random.default_rng(1)
    u, v, w = wall_bounded_velocity(g, rng, dtype)
    ut, vt, wt = [np.zeros_like(u) for _ in range(3)]
    P = O.Pres4(g); p = g.field(); dt = 0.5
    div0 = float(P.divergence(u, v, w))
    P.exec(p, u, v, w, ut, vt, wt, dt)
    un, vn, wn = u + dtype(dt)*ut, v + dtype(dt)*vt, w + dtype(dt)*wt
    ks, ke = g.kstart, g.kend
    wn[ks-1] = -wn[ks+1]; wn[ke+1] = -wn[ke-1]
    for a in (un, vn, wn):
        O.boundary_cyclic(g, a)
    assert div0 > 1.
    assert float(P.divergence(un, vn, wn)) <= EPS[dtype]*div0
    # ghost levels of p: zero gradient, two deep (src/pres_4.cxx:507-528)
    I = (slice(g.jstart, g.jend), slice(g.istart, g.iend))
    assert np.array_equal(p[ks-1][I], p[ks][I]) and np.array_equal(p[ks-2][I], p[ks+1][I])
    assert np.array_equal(p[ke][I], p[ke-1][I]) and np.array_equal(p[ke+1][I], p[ke-2][I])


def test_hdma_solves_its_band_system():
    g = O.Grid(8, 4, 9, 1., 1., 1., 3, 3, 3, np.float64, order=4)
    P = O.Pres4(g)
    rng = np.random.default_rng(0)
    n = g.kmax + 4
    bands = [rng.standard_normal((n, 3)) * 0.1 for _ in range(7)]
    bands[3] = 2. + rng.random((n, 3))                     # diagonally dominant
    rhs = rng.standard_normal((n, 3))
    A = np.zeros((3, n, n))
    for c in range(3):
        for r in range(n):
            for b in range(7):
                q = r + b - 3
                if 0 <= q < n:
                    A[c, r, q] = bands[b][r, c]
    x = rhs.copy()
    P.hdma(*[b.copy() for b in bands], x)
    for c in range(3):
        assert np.allclose(A[c] @ x[:, c], rhs[:, c], rtol=0, atol=1e-12)
