"""
TEST INFRASTRUCTURE ONLY -- the parity oracle.  Nothing in the product path
(`microhh_b200/`) may import this module; only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` leg do.

A CPU restatement, in numpy, of MicroHH's dynamical-core hot path.  Every function
cites the reference file:line it follows.  Arrays are C-ordered `(kcells, jcells, icells)`
views of the reference's `ijk = i + j*icells + k*ijcells` layout, so `a[k, j, i]`
is the reference's `a[ijk]`.  Expression grouping follows the reference so that the
results are bit-identical to the reference's own CPU kernels compiled with
`-ffp-contract=off` (`oracle/_ref/libmhh_ref.so`; checked in tests/test_oracle_vs_ref.py
and pinned by the golden vectors in tests/golden/).

Parity status: PINNED, bit for bit, against the reference's own compiled code (tests/test_oracle_vs_ref.py):
  * free-function kernels: advec_2i5 / advec_2 / advec_4 / advec_4m / diff_smag2 / diff_2 / diff_4 / thermo_dry /
    boundary (cyclic, 2nd- and 4th-order vertical ghost cells) / rk3 / tdma / Monin-Obukhov surface solver / buffer / force;
  * class-member code, run on stand-in objects (oracle/ref/ref_fake_pres.h, ref_grid.cpp): Grid::init + calculate (all
    metrics incl. dzi4 / dzhi4), FFT::init / load / exec_forward / exec_backward (plans, slice loops, normalisation),
    Pres_2 and Pres_4 set_values / input / solve (+ hdma) / output / calc_divergence, and full RK3 steps through them.
ONE thing is "parity unpinned": the 1-D transform inside FFTW3 itself (a system package that is not in the reference
tree nor in this image; call sites src/fft.cxx:145-155).  It is restated from FFTW's published R2HC / HC2R definition with
pocketfft (numpy / scipy), checked against the DFT definition and through the post-pressure divergence; the reference's
plans are executed by THIS transform in the tier-2 pin (oracle/ref/ref_fftw_shim.cpp), so last-bit differences between
FFTW's and pocketfft's butterflies are the only part of the path no test here can see.
"""
import numpy as np

# --------------------------------------------------------------------------------------
# Grid (reference src/grid.cxx:141-170 sizes, :245-304 metrics for swspatialorder=2)
# --------------------------------------------------------------------------------------
class Grid:
    def __init__(self, itot, jtot, ktot, xsize, ysize, zsize, igc, jgc, kgc, dtype=np.float64, z=None, order=2):
        TF = np.dtype(dtype).type
        self.TF = TF
        self.itot, self.jtot, self.ktot = itot, jtot, ktot
        self.imax, self.jmax, self.kmax = itot, jtot, ktot
        self.igc, self.jgc, self.kgc = igc, jgc, kgc
        self.icells, self.jcells, self.kcells = itot + 2*igc, jtot + 2*jgc, ktot + 2*kgc
        self.ijcells = self.icells*self.jcells
        self.ncells = self.ijcells*self.kcells
        self.istart, self.iend = igc, igc + itot
        self.jstart, self.jend = jgc, jgc + jtot
        self.kstart, self.kend = kgc, kgc + ktot
        self.xsize, self.ysize, self.zsize = TF(xsize), TF(ysize), TF(zsize)
        # src/grid.cxx:250-253
        self.dx = TF(self.xsize / itot)
        self.dy = TF(self.ysize / jtot)
        kc = self.kcells
        ks, ke = self.kstart, self.kend
        self.z = np.zeros(kc, TF); self.zh = np.zeros(kc, TF)
        self.dz = np.zeros(kc, TF); self.dzh = np.zeros(kc, TF)
        self.dzi = np.zeros(kc, TF); self.dzhi = np.zeros(kc, TF)
        if z is None:
            # uniform grid as cases/drycblles/drycblles_input.py:17-27 builds it
            dz = zsize / ktot
            z = np.linspace(0.5*dz, zsize - 0.5*dz, ktot)
        self.z[ks:ke] = np.asarray(z, TF)
        self.order = order
        if order == 4:
            assert kgc >= 3 and igc >= 3, "swspatialorder=4 uses three ghost cells (src/grid.cxx:87-92)"
            self._calculate_4th()
        else:
            self._calculate_2nd()

    def _calculate_4th(self):
        """src/grid.cxx:306-375.  Literals like `2.` / `(1./3.)` are double in the reference, so those lines are
        evaluated in double and narrowed on assignment; `ci0<TF>*z[..]` lines stay in TF."""
        TF = self.TF
        D = np.float64
        ks, ke, kc = self.kstart, self.kend, self.kcells
        z, zh, dz, dzh, dzi, dzhi = self.z, self.zh, self.dz, self.dzh, self.dzi, self.dzhi
        self.dzi4 = np.zeros(kc, TF); self.dzhi4 = np.zeros(kc, TF)
        dzi4, dzhi4 = self.dzi4, self.dzhi4
        dhuge = 1e30
        z[ks-1] = TF(-2.*D(z[ks]) + (1./3.)*D(z[ks+1]))
        z[ks-2] = TF(-9.*D(z[ks]) + 2.*D(z[ks+1]))
        z[ke] = TF((8./3.)*D(self.zsize) - 2.*D(z[ke-1]) + (1./3.)*D(z[ke-2]))
        z[ke+1] = TF(8.*D(self.zsize) - 9.*D(z[ke-1]) + 2.*D(z[ke-2]))
        z[ks-3] = TF(dhuge); z[ke+2] = TF(dhuge)
        ci = [TF(c) for c in CI]; bi = [TF(c) for c in BI]; ti = [TF(c) for c in TI]
        cg = [TF(c) for c in CG]; bg = [TF(c) for c in BG]; tg = [TF(c) for c in TG]
        w4 = lambda c, a, k0: c[0]*a[k0] + c[1]*a[k0+1] + c[2]*a[k0+2] + c[3]*a[k0+3]
        zh[ks] = TF(0.)
        for k in range(ks+1, ke):
            zh[k] = w4(ci, z, k-2)
        zh[ke] = self.zsize
        zh[ks-1] = w4(bi, z, ks-2)
        zh[ke+1] = w4(ti, z, ke-2)
        with np.errstate(over="ignore", invalid="ignore"):
            for k in range(1, kc):
                dzh[k] = z[k] - z[k-1]
                dzhi[k] = TF(1./D(dzh[k]))
            dzh[ks-3] = dzh[ks+3]; dzhi[ks-3] = dzhi[ks+3]
            for k in range(1, kc-1):
                dz[k] = zh[k+1] - zh[k]
                dzi[k] = TF(1./D(dz[k])) if dz[k] != 0 else TF(np.inf)
        dz[ks-3] = dz[ks+2]; dzi[ks-3] = dzi[ks+2]
        dz[ke+2] = dz[ke-3]; dzi[ke+2] = dzi[ke-3]
        for k in range(ks, ke):
            dzi4[k] = TF(1./D(w4(cg, zh, k-1)))
            dzhi4[k] = TF(1./D(w4(cg, z, k-2)))
        dzhi4[ke] = TF(1./D(w4(cg, z, ke-2)))
        dzi4[ks-1] = TF(1./D(w4(bg, zh, ks-1)))
        dzhi4[ks-1] = TF(1./D(w4(bg, z, ks-2)))
        dzi4[ke] = TF(1./D(w4(tg, zh, ke-2)))
        dzhi4[ke+1] = TF(1./D(w4(tg, z, ke-2)))
        self.dzhi4bot = TF(1./D(w4(bg, z, ks-1)))
        self.dzhi4top = TF(1./D(w4(tg, z, ke-3)))
        for k in (ks-2, ks-3, ke+1, ke+2):
            dzi4[k] = TF(dhuge)

    def _calculate_2nd(self):
        """src/grid.cxx:274-304"""
        TF = self.TF
        ks, ke, kc = self.kstart, self.kend, self.kcells
        z, zh, dz, dzh, dzi, dzhi = self.z, self.zh, self.dz, self.dzh, self.dzi, self.dzhi
        z[ks-1] = -z[ks]
        z[ke] = TF(2.)*self.zsize - z[ke-1]
        for k in range(ks+1, ke):
            zh[k] = TF(0.5)*(z[k-1] + z[k])
        zh[ks] = TF(0.)
        zh[ke] = self.zsize
        for k in range(1, kc):
            dzh[k] = z[k] - z[k-1]
            dzhi[k] = TF(1.)/dzh[k]
        dzh[ks-1] = dzh[ks+1]
        dzhi[ks-1] = dzhi[ks+1]
        for k in range(1, kc-1):
            dz[k] = zh[k+1] - zh[k]
            with np.errstate(divide="ignore"):      # kgc = 2: dz = 0 at the outer ghost level, as in the reference (never read)
                dzi[k] = TF(1.)/dz[k]
        dz[ks-1] = dz[ks]; dzi[ks-1] = dzi[ks]
        dz[ke] = dz[ke-1]; dzi[ke] = dzi[ke-1]

    def field(self, fill=0.):
        return np.full((self.kcells, self.jcells, self.icells), fill, self.TF)

    def field2d(self, fill=0.):
        return np.full((self.jcells, self.icells), fill, self.TF)


def _S(g, a, dk=0, dj=0, di=0, k0=None, k1=None):
    """Interior view of `a` shifted by (dk,dj,di); k-range defaults to kstart..kend."""
    k0 = g.kstart if k0 is None else k0
    k1 = g.kend if k1 is None else k1
    return a[k0+dk:k1+dk, g.jstart+dj:g.jend+dj, g.istart+di:g.iend+di]


def _K(g, v, dk=0, k0=None, k1=None):
    """1-D profile slice broadcastable over an interior view."""
    k0 = g.kstart if k0 is None else k0
    k1 = g.kend if k1 is None else k1
    return v[k0+dk:k1+dk, None, None]


# --------------------------------------------------------------------------------------
# Finite-difference helpers (reference include/finite_difference.h:33-158)
# --------------------------------------------------------------------------------------
# 4th-order weights (include/finite_difference.h:58-93)
CI = (-1./16., 9./16., 9./16., -1./16.)
BI = (5./16., 15./16., -5./16., 1./16.)
TI = (1./16., -5./16., 15./16., 5./16.)
CG = (1./24., -27./24., 27./24., -1./24.)
BG = (-23./24., 21./24., 3./24., -1./24.)
TG = (1./24., -3./24., -21./24., 23./24.)
CDG = (-1460./576., 783./576., -54./576., 1./576.)
def interp2(a, b):
    return a.dtype.type(0.5)*(a + b)

def interp4_ws(a, b, c, d):
    T = a.dtype.type
    return T(7./12.)*(b + c) - T(1./12.)*(a + d)

def interp3_ws(a, b, c, d):
    T = a.dtype.type
    return T(3./12.)*(c - b) - T(1./12.)*(d - a)

def interp6_ws(a, b, c, d, e, f):
    T = a.dtype.type
    return T(37./60.)*(c + d) - T(8./60.)*(b + e) + T(1./60.)*(a + f)

def interp5_ws(a, b, c, d, e, f):
    T = a.dtype.type
    return T(10./60.)*(d - c) - T(5./60.)*(e - b) + T(1./60.)*(f - a)


# --------------------------------------------------------------------------------------
# Boundary_cyclic (reference src/boundary_cyclic.cxx:369-443, 445-507)
# --------------------------------------------------------------------------------------
EDGE_EW, EDGE_NS, EDGE_BOTH = 0, 1, 2

def boundary_cyclic(g, a, edge=EDGE_BOTH):
    igc, jgc = g.igc, g.jgc
    if edge in (EDGE_EW, EDGE_BOTH):
        a[:, :, 0:igc] = a[:, :, g.iend-igc:g.iend]
        a[:, :, g.iend:g.iend+igc] = a[:, :, g.istart:g.istart+igc]
    if edge in (EDGE_NS, EDGE_BOTH):
        if g.jtot > 1:
            a[:, 0:jgc, :] = a[:, g.jend-jgc:g.jend, :]
            a[:, g.jend:g.jend+jgc, :] = a[:, g.jstart:g.jstart+jgc, :]
        else:
            ref = a[g.kstart:g.kend, g.jstart:g.jstart+1, :]
            a[g.kstart:g.kend, 0:jgc, :] = ref
            a[g.kstart:g.kend, g.jend:g.jend+jgc, :] = ref

def boundary_cyclic_2d(g, a):
    igc, jgc = g.igc, g.jgc
    a[:, 0:igc] = a[:, g.iend-igc:g.iend]
    a[:, g.iend:g.iend+igc] = a[:, g.istart:g.istart+igc]
    if g.jtot > 1:
        a[0:jgc, :] = a[g.jend-jgc:g.jend, :]
        a[g.jend:g.jend+jgc, :] = a[g.jstart:g.jstart+jgc, :]
    else:
        a[0:jgc, :] = a[g.jstart:g.jstart+1, :]
        a[g.jend:g.jend+jgc, :] = a[g.jstart:g.jstart+1, :]


# --------------------------------------------------------------------------------------
# Vertical ghost cells, 2nd order (reference src/boundary.cxx:700-772)
# --------------------------------------------------------------------------------------
BC_DIRICHLET, BC_NEUMANN = 0, 1

def ghost_cells_bot_2nd(g, a, bc, abot, agradbot):
    ks = g.kstart
    if bc == BC_DIRICHLET:
        a[ks-1] = g.TF(2.)*abot - a[ks]
    else:
        a[ks-1] = -agradbot*g.dzh[ks] + a[ks]

def ghost_cells_top_2nd(g, a, bc, atop, agradtop):
    ke = g.kend
    if bc == BC_DIRICHLET:
        a[ke] = g.TF(2.)*atop - a[ke-1]
    else:
        a[ke] = agradtop*g.dzh[ke] + a[ke-1]


# --------------------------------------------------------------------------------------
# 4th-order vertical ghost cells (reference src/boundary.cxx:776-922): two ghost levels from the wall value
# (Dirichlet) or the wall gradient (Neumann / flux); w: no-penetration, "normal" (one level, 3rd-order extrapolation of
# the zero-divergence condition) and "conservation" (two levels mirrored with a sign flip) types.
# --------------------------------------------------------------------------------------
def _grad4(a, b, c, d, TF):
    """include/finite_difference.h:127-131: -cg0*(d-a) - cg1*(c-b)"""
    return -TF(CG[0])*(d - a) - TF(CG[1])*(c - b)


def ghost_cells_bot_4th(g, a, bc, abot, agradbot):
    TF = g.TF; ks = g.kstart
    if bc == BC_DIRICHLET:
        a[ks-1] = TF(8./3.)*abot - TF(2.)*a[ks] + TF(1./3.)*a[ks+1]
        a[ks-2] = TF(8.)*abot - TF(9.)*a[ks] + TF(2.)*a[ks+1]
    elif bc == BC_NEUMANN:
        gr = _grad4(g.z[ks-2], g.z[ks-1], g.z[ks], g.z[ks+1], TF)
        a[ks-1] = TF(-1.)*gr*agradbot + a[ks]
        a[ks-2] = TF(-3.)*gr*agradbot + a[ks+1]


def ghost_cells_top_4th(g, a, bc, atop, agradtop):
    TF = g.TF; ke = g.kend
    if bc == BC_DIRICHLET:
        a[ke] = TF(8./3.)*atop - TF(2.)*a[ke-1] + TF(1./3.)*a[ke-2]
        a[ke+1] = TF(8.)*atop - TF(9.)*a[ke-1] + TF(2.)*a[ke-2]
    elif bc == BC_NEUMANN:
        gr = _grad4(g.z[ke-2], g.z[ke-1], g.z[ke], g.z[ke+1], TF)
        a[ke] = TF(1.)*gr*agradtop + a[ke-1]
        a[ke+1] = TF(3.)*gr*agradtop + a[ke-2]


def ghost_cells_w_4th(g, w, conservation):
    TF = g.TF; ks, ke = g.kstart, g.kend
    if conservation:
        w[ks-1] = -w[ks+1]; w[ks-2] = -w[ks+2]
        w[ke+1] = -w[ke-1]; w[ke+2] = -w[ke-2]
    else:
        w[ks-1] = TF(-6.)*w[ks+1] + TF(4.)*w[ks+2] - w[ks+3]
        w[ke+1] = TF(-6.)*w[ke-1] + TF(4.)*w[ke-2] - w[ke-3]


# --------------------------------------------------------------------------------------
# Advec_2i5 (reference src/advec_2i5.cxx:151-728)
# --------------------------------------------------------------------------------------
def _advec_2i5_vertical(g, at, a, wface, rho_f, rho_c, dzx, lo, hi):
    """Shared vertical structure of advec_u/v/s (faces k, k+1 of cell k; weights
    rhorefh[k+1], rhorefh[k] / rhoref[k] * dzi[k]) -- src/advec_2i5.cxx:203-299.
    wface(dk, k0, k1) returns the advecting velocity on face k+dk for cells k0..k1."""
    ks, ke = g.kstart, g.kend
    A = lambda dk, k0, k1: _S(g, a, dk, 0, 0, k0, k1)
    R = lambda v, dk, k0, k1: _K(g, v, dk, k0, k1)

    # interior, full 5/6th order (:204-217)
    k0, k1 = ks+3, ke-3
    if k1 > k0:
        wt_, wb_ = wface(1, k0, k1), wface(0, k0, k1)
        _S(g, at, 0, 0, 0, k0, k1)[...] += (
            - ( R(rho_f, 1, k0, k1) * wt_ * interp6_ws(A(-2,k0,k1), A(-1,k0,k1), A(0,k0,k1), A(1,k0,k1), A(2,k0,k1), A(3,k0,k1))
              - R(rho_f, 0, k0, k1) * wb_ * interp6_ws(A(-3,k0,k1), A(-2,k0,k1), A(-1,k0,k1), A(0,k0,k1), A(1,k0,k1), A(2,k0,k1)) ) / R(rho_c, 0, k0, k1) * R(dzx, 0, k0, k1)
            + ( R(rho_f, 1, k0, k1) * np.abs(wt_) * interp5_ws(A(-2,k0,k1), A(-1,k0,k1), A(0,k0,k1), A(1,k0,k1), A(2,k0,k1), A(3,k0,k1))
              - R(rho_f, 0, k0, k1) * np.abs(wb_) * interp5_ws(A(-3,k0,k1), A(-2,k0,k1), A(-1,k0,k1), A(0,k0,k1), A(1,k0,k1), A(2,k0,k1)) ) / R(rho_c, 0, k0, k1) * R(dzx, 0, k0, k1) )

    # k = kstart (:220-229)
    k0, k1 = ks, ks+1
    wt_ = wface(1, k0, k1)
    _S(g, at, 0, 0, 0, k0, k1)[...] += (
        - ( R(rho_f, 1, k0, k1) * wt_ * interp2(A(0,k0,k1), A(1,k0,k1)) ) / R(rho_c, 0, k0, k1) * R(dzx, 0, k0, k1) )

    # k = kstart+1 (:231-243)
    k0, k1 = ks+1, ks+2
    wt_, wb_ = wface(1, k0, k1), wface(0, k0, k1)
    _S(g, at, 0, 0, 0, k0, k1)[...] += (
        - ( R(rho_f, 1, k0, k1) * wt_ * interp4_ws(A(-1,k0,k1), A(0,k0,k1), A(1,k0,k1), A(2,k0,k1))
          - R(rho_f, 0, k0, k1) * wb_ * interp2(A(-1,k0,k1), A(0,k0,k1)) ) / R(rho_c, 0, k0, k1) * R(dzx, 0, k0, k1)
        + ( R(rho_f, 1, k0, k1) * np.abs(wt_) * interp3_ws(A(-1,k0,k1), A(0,k0,k1), A(1,k0,k1), A(2,k0,k1)) ) / R(rho_c, 0, k0, k1) * R(dzx, 0, k0, k1) )

    # k = kstart+2 (:246-259)
    k0, k1 = ks+2, ks+3
    wt_, wb_ = wface(1, k0, k1), wface(0, k0, k1)
    _S(g, at, 0, 0, 0, k0, k1)[...] += (
        - ( R(rho_f, 1, k0, k1) * wt_ * interp6_ws(A(-2,k0,k1), A(-1,k0,k1), A(0,k0,k1), A(1,k0,k1), A(2,k0,k1), A(3,k0,k1))
          - R(rho_f, 0, k0, k1) * wb_ * interp4_ws(A(-2,k0,k1), A(-1,k0,k1), A(0,k0,k1), A(1,k0,k1)) ) / R(rho_c, 0, k0, k1) * R(dzx, 0, k0, k1)
        + ( R(rho_f, 1, k0, k1) * np.abs(wt_) * interp5_ws(A(-2,k0,k1), A(-1,k0,k1), A(0,k0,k1), A(1,k0,k1), A(2,k0,k1), A(3,k0,k1))
          - R(rho_f, 0, k0, k1) * np.abs(wb_) * interp3_ws(A(-2,k0,k1), A(-1,k0,k1), A(0,k0,k1), A(1,k0,k1)) ) / R(rho_c, 0, k0, k1) * R(dzx, 0, k0, k1) )

    # k = kend-3 (:261-274)
    k0, k1 = ke-3, ke-2
    wt_, wb_ = wface(1, k0, k1), wface(0, k0, k1)
    _S(g, at, 0, 0, 0, k0, k1)[...] += (
        - ( R(rho_f, 1, k0, k1) * wt_ * interp4_ws(A(-1,k0,k1), A(0,k0,k1), A(1,k0,k1), A(2,k0,k1))
          - R(rho_f, 0, k0, k1) * wb_ * interp6_ws(A(-3,k0,k1), A(-2,k0,k1), A(-1,k0,k1), A(0,k0,k1), A(1,k0,k1), A(2,k0,k1)) ) / R(rho_c, 0, k0, k1) * R(dzx, 0, k0, k1)
        + ( R(rho_f, 1, k0, k1) * np.abs(wt_) * interp3_ws(A(-1,k0,k1), A(0,k0,k1), A(1,k0,k1), A(2,k0,k1))
          - R(rho_f, 0, k0, k1) * np.abs(wb_) * interp5_ws(A(-3,k0,k1), A(-2,k0,k1), A(-1,k0,k1), A(0,k0,k1), A(1,k0,k1), A(2,k0,k1)) ) / R(rho_c, 0, k0, k1) * R(dzx, 0, k0, k1) )

    # k = kend-2 (:276-288)
    k0, k1 = ke-2, ke-1
    wt_, wb_ = wface(1, k0, k1), wface(0, k0, k1)
    _S(g, at, 0, 0, 0, k0, k1)[...] += (
        - ( R(rho_f, 1, k0, k1) * wt_ * interp2(A(0,k0,k1), A(1,k0,k1))
          - R(rho_f, 0, k0, k1) * wb_ * interp4_ws(A(-2,k0,k1), A(-1,k0,k1), A(0,k0,k1), A(1,k0,k1)) ) / R(rho_c, 0, k0, k1) * R(dzx, 0, k0, k1)
        - ( R(rho_f, 0, k0, k1) * np.abs(wb_) * interp3_ws(A(-2,k0,k1), A(-1,k0,k1), A(0,k0,k1), A(1,k0,k1)) ) / R(rho_c, 0, k0, k1) * R(dzx, 0, k0, k1) )

    # k = kend-1 (:290-299)
    k0, k1 = ke-1, ke
    wb_ = wface(0, k0, k1)
    _S(g, at, 0, 0, 0, k0, k1)[...] += (
        - ( -R(rho_f, 0, k0, k1) * wb_ * interp2(A(-1,k0,k1), A(0,k0,k1)) ) / R(rho_c, 0, k0, k1) * R(dzx, 0, k0, k1) )


def advec_2i5_u(g, ut, u, v, w, rhoref, rhorefh):
    """src/advec_2i5.cxx:151-300"""
    TF = g.TF
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    U = lambda di=0, dj=0: _S(g, u, 0, dj, di)
    V = lambda di=0, dj=0: _S(g, v, 0, dj, di)
    _S(g, ut)[...] += (
        - ( interp2(U(0), U(1)) * interp6_ws(U(-2), U(-1), U(0), U(1), U(2), U(3))
          - interp2(U(-1), U(0)) * interp6_ws(U(-3), U(-2), U(-1), U(0), U(1), U(2)) ) * dxi
        + ( np.abs(interp2(U(0), U(1))) * interp5_ws(U(-2), U(-1), U(0), U(1), U(2), U(3))
          - np.abs(interp2(U(-1), U(0))) * interp5_ws(U(-3), U(-2), U(-1), U(0), U(1), U(2)) ) * dxi
        - ( interp2(V(-1, 1), V(0, 1)) * interp6_ws(U(0,-2), U(0,-1), U(0,0), U(0,1), U(0,2), U(0,3))
          - interp2(V(-1, 0), V(0, 0)) * interp6_ws(U(0,-3), U(0,-2), U(0,-1), U(0,0), U(0,1), U(0,2)) ) * dyi
        + ( np.abs(interp2(V(-1, 1), V(0, 1))) * interp5_ws(U(0,-2), U(0,-1), U(0,0), U(0,1), U(0,2), U(0,3))
          - np.abs(interp2(V(-1, 0), V(0, 0))) * interp5_ws(U(0,-3), U(0,-2), U(0,-1), U(0,0), U(0,1), U(0,2)) ) * dyi )
    wface = lambda dk, k0, k1: interp2(_S(g, w, dk, 0, -1, k0, k1), _S(g, w, dk, 0, 0, k0, k1))
    _advec_2i5_vertical(g, ut, u, wface, rhorefh, rhoref, g.dzi, 0, 0)


def advec_2i5_v(g, vt, u, v, w, rhoref, rhorefh):
    """src/advec_2i5.cxx:303-450"""
    TF = g.TF
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    U = lambda di=0, dj=0: _S(g, u, 0, dj, di)
    V = lambda di=0, dj=0: _S(g, v, 0, dj, di)
    _S(g, vt)[...] += (
        - ( interp2(U(1,-1), U(1,0)) * interp6_ws(V(-2), V(-1), V(0), V(1), V(2), V(3))
          - interp2(U(0,-1), U(0,0)) * interp6_ws(V(-3), V(-2), V(-1), V(0), V(1), V(2)) ) * dxi
        + ( np.abs(interp2(U(1,-1), U(1,0))) * interp5_ws(V(-2), V(-1), V(0), V(1), V(2), V(3))
          - np.abs(interp2(U(0,-1), U(0,0))) * interp5_ws(V(-3), V(-2), V(-1), V(0), V(1), V(2)) ) * dxi
        - ( interp2(V(0,0), V(0,1)) * interp6_ws(V(0,-2), V(0,-1), V(0,0), V(0,1), V(0,2), V(0,3))
          - interp2(V(0,-1), V(0,0)) * interp6_ws(V(0,-3), V(0,-2), V(0,-1), V(0,0), V(0,1), V(0,2)) ) * dyi
        + ( np.abs(interp2(V(0,0), V(0,1))) * interp5_ws(V(0,-2), V(0,-1), V(0,0), V(0,1), V(0,2), V(0,3))
          - np.abs(interp2(V(0,-1), V(0,0))) * interp5_ws(V(0,-3), V(0,-2), V(0,-1), V(0,0), V(0,1), V(0,2)) ) * dyi )
    wface = lambda dk, k0, k1: interp2(_S(g, w, dk, -1, 0, k0, k1), _S(g, w, dk, 0, 0, k0, k1))
    _advec_2i5_vertical(g, vt, v, wface, rhorefh, rhoref, g.dzi, 0, 0)


def advec_2i5_s(g, st, s, u, v, w, rhoref, rhorefh):
    """src/advec_2i5.cxx:582-728"""
    TF = g.TF
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    S = lambda di=0, dj=0: _S(g, s, 0, dj, di)
    U = lambda di=0, dj=0: _S(g, u, 0, dj, di)
    V = lambda di=0, dj=0: _S(g, v, 0, dj, di)
    _S(g, st)[...] += (
        - ( U(1) * interp6_ws(S(-2), S(-1), S(0), S(1), S(2), S(3))
          - U(0) * interp6_ws(S(-3), S(-2), S(-1), S(0), S(1), S(2)) ) * dxi
        + ( np.abs(U(1)) * interp5_ws(S(-2), S(-1), S(0), S(1), S(2), S(3))
          - np.abs(U(0)) * interp5_ws(S(-3), S(-2), S(-1), S(0), S(1), S(2)) ) * dxi
        - ( V(0,1) * interp6_ws(S(0,-2), S(0,-1), S(0,0), S(0,1), S(0,2), S(0,3))
          - V(0,0) * interp6_ws(S(0,-3), S(0,-2), S(0,-1), S(0,0), S(0,1), S(0,2)) ) * dyi
        + ( np.abs(V(0,1)) * interp5_ws(S(0,-2), S(0,-1), S(0,0), S(0,1), S(0,2), S(0,3))
          - np.abs(V(0,0)) * interp5_ws(S(0,-3), S(0,-2), S(0,-1), S(0,0), S(0,1), S(0,2)) ) * dyi )
    wface = lambda dk, k0, k1: _S(g, w, dk, 0, 0, k0, k1)
    _advec_2i5_vertical(g, st, s, wface, rhorefh, rhoref, g.dzi, 0, 0)
    # NOTE: kend-2 abs term in advec_s is written "+ ( -rhorefh[k]*abs(w)*interp3 )" (:715) which is
    # bit-identical to the "- ( rhorefh[k]*abs(w)*interp3 )" form of advec_u/v used above.


# --------------------------------------------------------------------------------------
# Flux-limited scalar advection (reference include/advec_monotonic.h:28-202): Koren (1993) limiter, used by Advec_2i5 for
# the scalars in `fluxlimit_list` (src/advec_2i5.cxx:1046-1056).  Faces kstart / kend carry no flux; the first / last
# interior face falls back to first-order upwind on its wall side (flux_lim_bot / flux_lim_top).
# --------------------------------------------------------------------------------------
def _flux_lim(u, sm2, sm1, sp1, sp2, variant=0):
    """variant 0: flux_lim, 1: flux_lim_bot (u >= 0 -> u*sm1), 2: flux_lim_top (u < 0 -> u*sp1)"""
    TF = u.dtype.type
    eps = np.finfo(TF).eps
    def limited(a2, a1, b1):
        # u*(a1 + 0.5*phi*(a1 - a2)), phi from two_r = 2*(b1 - a1)/denom
        d = a1 - a2
        denom = (np.copysign(1., d) * np.maximum(np.abs(d), eps)).astype(TF)
        two_r = TF(2.)*(b1 - a1)/denom
        phi = np.maximum(TF(0.), np.minimum(two_r, np.minimum(TF(1./3.)*(TF(1.) + two_r), TF(2.))))
        return u*(a1 + TF(0.5)*phi*(a1 - a2))
    pos = u*sm1 if variant == 1 else limited(sm2, sm1, sp1)
    neg = u*sp1 if variant == 2 else limited(sp2, sp1, sm1)
    return np.where(u >= TF(0.), pos, neg)


def advec_s_lim(g, st, s, u, v, w, rhoref, rhorefh):
    """include/advec_monotonic.h:98-202"""
    TF = g.TF
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    ks, ke = g.kstart, g.kend
    def horiz(k0, k1):
        S = lambda dj=0, di=0: _S(g, s, 0, dj, di, k0, k1)
        U = lambda di=0: _S(g, u, 0, 0, di, k0, k1)
        V = lambda dj=0: _S(g, v, 0, dj, 0, k0, k1)
        return ( - ( _flux_lim(U(1), S(0, -1), S(), S(0, 1), S(0, 2)) - _flux_lim(U(0), S(0, -2), S(0, -1), S(), S(0, 1)) ) * dxi
                 - ( _flux_lim(V(1), S(-1), S(), S(1), S(2)) - _flux_lim(V(0), S(-2), S(-1), S(), S(1)) ) * dyi )
    def vface(k0, k1, up, variant):
        # flux through the upper (up=1) or lower (up=0) face of rows k0..k1
        Sk = lambda dk: _S(g, s, dk, 0, 0, k0, k1)
        return _K(g, rhorefh, up, k0, k1) * _flux_lim(_S(g, w, up, 0, 0, k0, k1), Sk(up-2), Sk(up-1), Sk(up), Sk(up+1), variant)
    def add(k0, k1, top, bot):
        if k1 <= k0:
            return
        vert = (top - bot) if (top is not None and bot is not None) else (top if bot is None else -bot)
        _S(g, st, 0, 0, 0, k0, k1)[...] += horiz(k0, k1) - (vert) / _K(g, rhoref, 0, k0, k1) * _K(g, g.dzi, 0, k0, k1)
    add(ks+2, ke-2, vface(ks+2, ke-2, 1, 0), vface(ks+2, ke-2, 0, 0))
    add(ks, ks+1, vface(ks, ks+1, 1, 1), None)
    add(ks+1, ks+2, vface(ks+1, ks+2, 1, 0), vface(ks+1, ks+2, 0, 1))
    add(ke-2, ke-1, vface(ke-2, ke-1, 1, 2), vface(ke-2, ke-1, 0, 0))
    add(ke-1, ke, None, vface(ke-1, ke, 0, 2))


def advec_2i5_w(g, wt, u, v, w, rhoref, rhorefh):
    """src/advec_2i5.cxx:453-579"""
    TF = g.TF
    ks, ke = g.kstart, g.kend
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    k0, k1 = ks+1, ke
    W = lambda di=0, dj=0: _S(g, w, 0, dj, di, k0, k1)
    U = lambda di=0, dj=0, dk=0: _S(g, u, dk, dj, di, k0, k1)
    V = lambda di=0, dj=0, dk=0: _S(g, v, dk, dj, di, k0, k1)
    _S(g, wt, 0, 0, 0, k0, k1)[...] += (
        - ( interp2(U(1,0,-1), U(1,0,0)) * interp6_ws(W(-2), W(-1), W(0), W(1), W(2), W(3))
          - interp2(U(0,0,-1), U(0,0,0)) * interp6_ws(W(-3), W(-2), W(-1), W(0), W(1), W(2)) ) * dxi
        + ( np.abs(interp2(U(1,0,-1), U(1,0,0))) * interp5_ws(W(-2), W(-1), W(0), W(1), W(2), W(3))
          - np.abs(interp2(U(0,0,-1), U(0,0,0))) * interp5_ws(W(-3), W(-2), W(-1), W(0), W(1), W(2)) ) * dxi
        - ( interp2(V(0,1,-1), V(0,1,0)) * interp6_ws(W(0,-2), W(0,-1), W(0,0), W(0,1), W(0,2), W(0,3))
          - interp2(V(0,0,-1), V(0,0,0)) * interp6_ws(W(0,-3), W(0,-2), W(0,-1), W(0,0), W(0,1), W(0,2)) ) * dyi
        + ( np.abs(interp2(V(0,1,-1), V(0,1,0))) * interp5_ws(W(0,-2), W(0,-1), W(0,0), W(0,1), W(0,2), W(0,3))
          - np.abs(interp2(V(0,0,-1), V(0,0,0))) * interp5_ws(W(0,-3), W(0,-2), W(0,-1), W(0,0), W(0,1), W(0,2)) ) * dyi )

    A = lambda dk, a0, a1: _S(g, w, dk, 0, 0, a0, a1)
    R = lambda vv, dk, a0, a1: _K(g, vv, dk, a0, a1)
    dzhi = g.dzhi

    # interior (:506-519)
    a0, a1 = ks+3, ke-2
    if a1 > a0:
        wt_ = interp2(A(0,a0,a1), A(1,a0,a1)); wb_ = interp2(A(-1,a0,a1), A(0,a0,a1))
        _S(g, wt, 0, 0, 0, a0, a1)[...] += (
            - ( R(rhoref, 0, a0, a1) * wt_ * interp6_ws(A(-2,a0,a1), A(-1,a0,a1), A(0,a0,a1), A(1,a0,a1), A(2,a0,a1), A(3,a0,a1))
              - R(rhoref,-1, a0, a1) * wb_ * interp6_ws(A(-3,a0,a1), A(-2,a0,a1), A(-1,a0,a1), A(0,a0,a1), A(1,a0,a1), A(2,a0,a1)) ) / R(rhorefh, 0, a0, a1) * R(dzhi, 0, a0, a1)
            + ( R(rhoref, 0, a0, a1) * np.abs(wt_) * interp5_ws(A(-2,a0,a1), A(-1,a0,a1), A(0,a0,a1), A(1,a0,a1), A(2,a0,a1), A(3,a0,a1))
              - R(rhoref,-1, a0, a1) * np.abs(wb_) * interp5_ws(A(-3,a0,a1), A(-2,a0,a1), A(-1,a0,a1), A(0,a0,a1), A(1,a0,a1), A(2,a0,a1)) ) / R(rhorefh, 0, a0, a1) * R(dzhi, 0, a0, a1) )

    # k = kstart+1 (:522-534)
    a0, a1 = ks+1, ks+2
    wt_ = interp2(A(0,a0,a1), A(1,a0,a1)); wb_ = interp2(A(-1,a0,a1), A(0,a0,a1))
    _S(g, wt, 0, 0, 0, a0, a1)[...] += (
        - ( R(rhoref, 0, a0, a1) * wt_ * interp4_ws(A(-1,a0,a1), A(0,a0,a1), A(1,a0,a1), A(2,a0,a1))
          - R(rhoref,-1, a0, a1) * wb_ * interp2(A(-1,a0,a1), A(0,a0,a1)) ) / R(rhorefh, 0, a0, a1) * R(dzhi, 0, a0, a1)
        + ( R(rhoref, 0, a0, a1) * np.abs(wt_) * interp3_ws(A(-1,a0,a1), A(0,a0,a1), A(1,a0,a1), A(2,a0,a1)) ) / R(rhorefh, 0, a0, a1) * R(dzhi, 0, a0, a1) )

    # k = kstart+2 (:536-549)
    a0, a1 = ks+2, ks+3
    wt_ = interp2(A(0,a0,a1), A(1,a0,a1)); wb_ = interp2(A(-1,a0,a1), A(0,a0,a1))
    _S(g, wt, 0, 0, 0, a0, a1)[...] += (
        - ( R(rhoref, 0, a0, a1) * wt_ * interp6_ws(A(-2,a0,a1), A(-1,a0,a1), A(0,a0,a1), A(1,a0,a1), A(2,a0,a1), A(3,a0,a1))
          - R(rhoref,-1, a0, a1) * wb_ * interp4_ws(A(-2,a0,a1), A(-1,a0,a1), A(0,a0,a1), A(1,a0,a1)) ) / R(rhorefh, 0, a0, a1) * R(dzhi, 0, a0, a1)
        + ( R(rhoref, 0, a0, a1) * np.abs(wt_) * interp5_ws(A(-2,a0,a1), A(-1,a0,a1), A(0,a0,a1), A(1,a0,a1), A(2,a0,a1), A(3,a0,a1))
          - R(rhoref,-1, a0, a1) * np.abs(wb_) * interp3_ws(A(-2,a0,a1), A(-1,a0,a1), A(0,a0,a1), A(1,a0,a1)) ) / R(rhorefh, 0, a0, a1) * R(dzhi, 0, a0, a1) )

    # k = kend-2 (:551-564)
    a0, a1 = ke-2, ke-1
    wt_ = interp2(A(0,a0,a1), A(1,a0,a1)); wb_ = interp2(A(-1,a0,a1), A(0,a0,a1))
    _S(g, wt, 0, 0, 0, a0, a1)[...] += (
        - ( R(rhoref, 0, a0, a1) * wt_ * interp4_ws(A(-1,a0,a1), A(0,a0,a1), A(1,a0,a1), A(2,a0,a1))
          - R(rhoref,-1, a0, a1) * wb_ * interp6_ws(A(-3,a0,a1), A(-2,a0,a1), A(-1,a0,a1), A(0,a0,a1), A(1,a0,a1), A(2,a0,a1)) ) / R(rhorefh, 0, a0, a1) * R(dzhi, 0, a0, a1)
        + ( R(rhoref, 0, a0, a1) * np.abs(wt_) * interp3_ws(A(-1,a0,a1), A(0,a0,a1), A(1,a0,a1), A(2,a0,a1))
          - R(rhoref,-1, a0, a1) * np.abs(wb_) * interp5_ws(A(-3,a0,a1), A(-2,a0,a1), A(-1,a0,a1), A(0,a0,a1), A(1,a0,a1), A(2,a0,a1)) ) / R(rhorefh, 0, a0, a1) * R(dzhi, 0, a0, a1) )

    # k = kend-1 (:566-578)
    a0, a1 = ke-1, ke
    wt_ = interp2(A(0,a0,a1), A(1,a0,a1)); wb_ = interp2(A(-1,a0,a1), A(0,a0,a1))
    _S(g, wt, 0, 0, 0, a0, a1)[...] += (
        - ( R(rhoref, 0, a0, a1) * wt_ * interp2(A(0,a0,a1), A(1,a0,a1))
          - R(rhoref,-1, a0, a1) * wb_ * interp4_ws(A(-2,a0,a1), A(-1,a0,a1), A(0,a0,a1), A(1,a0,a1)) ) / R(rhorefh, 0, a0, a1) * R(dzhi, 0, a0, a1)
        - ( R(rhoref,-1, a0, a1) * np.abs(wb_) * interp3_ws(A(-2,a0,a1), A(-1,a0,a1), A(0,a0,a1), A(1,a0,a1)) ) / R(rhorefh, 0, a0, a1) * R(dzhi, 0, a0, a1) )


def advec_2i5_cfl(g, u, v, w, dt):
    """src/advec_2i5.cxx:60-148 (returns cfl*dt in TF, as the reference does)"""
    TF = g.TF
    ks, ke = g.kstart, g.kend
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    cfl = TF(0.)
    def horiz(k0, k1):
        U = lambda di: _S(g, u, 0, 0, di, k0, k1)
        V = lambda dj: _S(g, v, 0, dj, 0, k0, k1)
        return (np.abs(interp6_ws(U(-2), U(-1), U(0), U(1), U(2), U(3)))*dxi
              + np.abs(interp6_ws(V(-2), V(-1), V(0), V(1), V(2), V(3)))*dyi)
    W = lambda dk, k0, k1: _S(g, w, dk, 0, 0, k0, k1)
    parts = []
    for (k0, k1) in ((ks, ks+1), (ke-1, ke)):
        parts.append(horiz(k0, k1) + np.abs(interp2(W(0,k0,k1), W(1,k0,k1)))*_K(g, g.dzi, 0, k0, k1))
    for (k0, k1) in ((ks+1, ks+2), (ke-2, ke-1)):
        parts.append(horiz(k0, k1) + np.abs(interp4_ws(W(-1,k0,k1), W(0,k0,k1), W(1,k0,k1), W(2,k0,k1)))*_K(g, g.dzi, 0, k0, k1))
    k0, k1 = ks+2, ke-2
    if k1 > k0:
        parts.append(horiz(k0, k1) + np.abs(interp6_ws(W(-2,k0,k1), W(-1,k0,k1), W(0,k0,k1), W(1,k0,k1), W(2,k0,k1), W(3,k0,k1)))*_K(g, g.dzi, 0, k0, k1))
    for p_ in parts:
        cfl = max(cfl, TF(p_.max()))
    return TF(cfl*TF(dt))


# --------------------------------------------------------------------------------------
# Advec_2 (reference src/advec_2.cxx:48-202): plain 2nd-order flux form, +-1 stencil
# --------------------------------------------------------------------------------------
def advec_2_u(g, ut, u, v, w, rhoref, rhorefh):
    """src/advec_2.cxx:78-107"""
    TF = g.TF
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    U = lambda dk=0, dj=0, di=0: _S(g, u, dk, dj, di)
    V = lambda dk=0, dj=0, di=0: _S(g, v, dk, dj, di)
    W = lambda dk=0, dj=0, di=0: _S(g, w, dk, dj, di)
    _S(g, ut)[...] += (
        - ( interp2(U(), U(0,0,1)) * interp2(U(), U(0,0,1))
          - interp2(U(0,0,-1), U()) * interp2(U(0,0,-1), U()) ) * dxi
        - ( interp2(V(0,1,-1), V(0,1,0)) * interp2(U(), U(0,1,0))
          - interp2(V(0,0,-1), V()) * interp2(U(0,-1,0), U()) ) * dyi
        - ( _K(g, rhorefh, 1) * interp2(W(1,0,-1), W(1,0,0)) * interp2(U(), U(1,0,0))
          - _K(g, rhorefh, 0) * interp2(W(0,0,-1), W()) * interp2(U(-1,0,0), U()) ) / _K(g, rhoref) * _K(g, g.dzi) )


def advec_2_v(g, vt, u, v, w, rhoref, rhorefh):
    """src/advec_2.cxx:109-138"""
    TF = g.TF
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    U = lambda dk=0, dj=0, di=0: _S(g, u, dk, dj, di)
    V = lambda dk=0, dj=0, di=0: _S(g, v, dk, dj, di)
    W = lambda dk=0, dj=0, di=0: _S(g, w, dk, dj, di)
    _S(g, vt)[...] += (
        - ( interp2(U(0,-1,1), U(0,0,1)) * interp2(V(), V(0,0,1))
          - interp2(U(0,-1,0), U()) * interp2(V(0,0,-1), V()) ) * dxi
        - ( interp2(V(), V(0,1,0)) * interp2(V(), V(0,1,0))
          - interp2(V(0,-1,0), V()) * interp2(V(0,-1,0), V()) ) * dyi
        - ( _K(g, rhorefh, 1) * interp2(W(1,-1,0), W(1,0,0)) * interp2(V(), V(1,0,0))
          - _K(g, rhorefh, 0) * interp2(W(0,-1,0), W()) * interp2(V(-1,0,0), V()) ) / _K(g, rhoref) * _K(g, g.dzi) )


def advec_2_w(g, wt, u, v, w, rhoref, rhorefh):
    """src/advec_2.cxx:140-169 (k = kstart+1 .. kend-1)"""
    TF = g.TF
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    k0, k1 = g.kstart+1, g.kend
    U = lambda dk=0, dj=0, di=0: _S(g, u, dk, dj, di, k0, k1)
    V = lambda dk=0, dj=0, di=0: _S(g, v, dk, dj, di, k0, k1)
    W = lambda dk=0, dj=0, di=0: _S(g, w, dk, dj, di, k0, k1)
    _S(g, wt, 0, 0, 0, k0, k1)[...] += (
        - ( interp2(U(-1,0,1), U(0,0,1)) * interp2(W(), W(0,0,1))
          - interp2(U(-1,0,0), U()) * interp2(W(0,0,-1), W()) ) * dxi
        - ( interp2(V(-1,1,0), V(0,1,0)) * interp2(W(), W(0,1,0))
          - interp2(V(-1,0,0), V()) * interp2(W(0,-1,0), W()) ) * dyi
        - ( _K(g, rhoref, 0, k0, k1) * interp2(W(), W(1,0,0)) * interp2(W(), W(1,0,0))
          - _K(g, rhoref, -1, k0, k1) * interp2(W(-1,0,0), W()) * interp2(W(-1,0,0), W()) ) / _K(g, rhorefh, 0, k0, k1) * _K(g, g.dzhi, 0, k0, k1) )


def advec_2_s(g, st, s, u, v, w, rhoref, rhorefh):
    """src/advec_2.cxx:171-202"""
    TF = g.TF
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    S = lambda dk=0, dj=0, di=0: _S(g, s, dk, dj, di)
    _S(g, st)[...] += (
        - ( _S(g, u, 0, 0, 1) * interp2(S(), S(0,0,1))
          - _S(g, u) * interp2(S(0,0,-1), S()) ) * dxi
        - ( _S(g, v, 0, 1, 0) * interp2(S(), S(0,1,0))
          - _S(g, v) * interp2(S(0,-1,0), S()) ) * dyi
        - ( _K(g, rhorefh, 1) * _S(g, w, 1) * interp2(S(), S(1,0,0))
          - _K(g, rhorefh, 0) * _S(g, w) * interp2(S(-1,0,0), S()) ) / _K(g, rhoref) * _K(g, g.dzi) )


def advec_2_cfl(g, u, v, w, dt):
    """src/advec_2.cxx:50-76 (returns cfl*dt in TF)"""
    TF = g.TF
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    c = (np.abs(interp2(_S(g, u), _S(g, u, 0, 0, 1)))*dxi + np.abs(interp2(_S(g, v), _S(g, v, 0, 1, 0)))*dyi
         + np.abs(interp2(_S(g, w), _S(g, w, 1)))*_K(g, g.dzi))
    return TF(TF(c.max())*TF(dt))


# --------------------------------------------------------------------------------------
# Diff_2 (reference src/diff_2.cxx:38-86, dnmul :139-152).  dxidxi / dyidyi are `double` even in the
# single-precision build (:44-45), so the SP sum is formed in double and narrowed by the `+=`.
# --------------------------------------------------------------------------------------
def _diff_2(g, at, a, visc, k0, dz_up, dz_dn, dz_c):
    TF = g.TF
    k1 = g.kend
    # `const double dxidxi = 1/(dx*dx);` -- int 1 over a TF product: the DIVISION is done in TF, the result widened
    dxidxi = np.float64(TF(1.)/TF(g.dx*g.dx))
    dyidyi = np.float64(TF(1.)/TF(g.dy*g.dy))
    A = lambda dk=0, dj=0, di=0: _S(g, a, dk, dj, di, k0, k1)
    lap = ( ( (A(0,0,1) - A()) - (A() - A(0,0,-1)) ) * dxidxi
          + ( (A(0,1,0) - A()) - (A() - A(0,-1,0)) ) * dyidyi
          + ( (A(1,0,0) - A()) * dz_up - (A() - A(-1,0,0)) * dz_dn ) * dz_c )
    tgt = _S(g, at, 0, 0, 0, k0, k1)
    tgt[...] = (tgt + TF(visc) * lap).astype(TF)


def diff_2_c(g, at, a, visc):
    """src/diff_2.cxx:38-61"""
    _diff_2(g, at, a, visc, g.kstart, _K(g, g.dzhi, 1), _K(g, g.dzhi, 0), _K(g, g.dzi))


def diff_2_w(g, wt, w, visc):
    """src/diff_2.cxx:63-86 (k = kstart+1 .. kend-1)"""
    k0, k1 = g.kstart+1, g.kend
    _diff_2(g, wt, w, visc, k0, _K(g, g.dzi, 0, k0, k1), _K(g, g.dzi, -1, k0, k1), _K(g, g.dzhi, 0, k0, k1))


def diff_2_dnmul(g, viscmax):
    """src/diff_2.cxx:139-152: max_k |viscmax (1/dx^2 + 1/dy^2 + 1/dz[k]^2)| (host side, once)"""
    TF = g.TF
    dz = g.dz[g.kstart:g.kend]
    return float(np.max(np.abs(TF(viscmax) * (1./np.float64(g.dx*g.dx) + 1./np.float64(g.dy*g.dy) + 1./(dz*dz).astype(np.float64)))))


# --------------------------------------------------------------------------------------
# Advec_4 (reference src/advec_4.cxx:50-487): 4th-order divergence (cg weights) of products of 4th-order
# interpolations (ci weights); the outermost vertical flux of the first / last row uses the one-sided
# bi / ti interpolation of the advected quantity.  Every direction is its own `-=` statement in the
# reference, so every direction is rounded into the tendency separately here as well.
# --------------------------------------------------------------------------------------
def _w4(c, a, b, cc, d, TF):
    """c0*a + c1*b + c2*c + c3*d, left to right, in TF"""
    return TF(c[0])*a + TF(c[1])*b + TF(c[2])*cc + TF(c[3])*d


def _advec4_rows(g, lo):
    """(k0, k1, bottom?, top?) row groups of a tendency whose first row is `lo`"""
    return ((lo, lo+1, True, False), (lo+1, g.kend-1, False, False), (g.kend-1, g.kend, False, True))


def _advec4_generic(g, at, q, vel_x, vel_y, vel_z, dzx, lo):
    """at -= d(vel_x q)/dx; at -= d(vel_y q)/dy (3-D only); at -= d(vel_z q)/dz.
    vel_d(m, k0, k1): advecting velocity at flux point m = 0..3 of direction d; q interpolated along d."""
    TF = g.TF
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    dim3 = g.jtot > 1
    for (k0, k1, bot, top) in _advec4_rows(g, lo):
        if k1 <= k0:
            continue
        Q = lambda dk=0, dj=0, di=0: _S(g, q, dk, dj, di, k0, k1)
        tgt = _S(g, at, 0, 0, 0, k0, k1)
        def div(vel, qi):
            return ( TF(CG[0])*(vel(0, k0, k1)*qi(0)) + TF(CG[1])*(vel(1, k0, k1)*qi(1))
                   + TF(CG[2])*(vel(2, k0, k1)*qi(2)) + TF(CG[3])*(vel(3, k0, k1)*qi(3)) )
        tgt[...] -= div(vel_x, lambda m: _w4(CI, Q(0, 0, m-3), Q(0, 0, m-2), Q(0, 0, m-1), Q(0, 0, m), TF)) * dxi
        if dim3:
            tgt[...] -= div(vel_y, lambda m: _w4(CI, Q(0, m-3), Q(0, m-2), Q(0, m-1), Q(0, m), TF)) * dyi
        def qz(m):
            if bot and m == 0:
                return _w4(BI, Q(-2), Q(-1), Q(0), Q(1), TF)
            if top and m == 3:
                return _w4(TI, Q(-1), Q(0), Q(1), Q(2), TF)
            return _w4(CI, Q(m-3), Q(m-2), Q(m-1), Q(m), TF)
        tgt[...] -= div(vel_z, qz) * _K(g, dzx, 0, k0, k1)


def advec_4_u(g, ut, u, v, w):
    """src/advec_4.cxx:88-186"""
    TF = g.TF
    I = lambda a, b, c, d: _w4(CI, a, b, c, d, TF)
    U = lambda k0, k1, dk=0, dj=0, di=0: _S(g, u, dk, dj, di, k0, k1)
    V = lambda k0, k1, dk=0, dj=0, di=0: _S(g, v, dk, dj, di, k0, k1)
    W = lambda k0, k1, dk=0, dj=0, di=0: _S(g, w, dk, dj, di, k0, k1)
    vx = lambda m, k0, k1: I(U(k0, k1, 0, 0, m-3), U(k0, k1, 0, 0, m-2), U(k0, k1, 0, 0, m-1), U(k0, k1, 0, 0, m))
    vy = lambda m, k0, k1: I(V(k0, k1, 0, m-1, -2), V(k0, k1, 0, m-1, -1), V(k0, k1, 0, m-1, 0), V(k0, k1, 0, m-1, 1))
    vz = lambda m, k0, k1: I(W(k0, k1, m-1, 0, -2), W(k0, k1, m-1, 0, -1), W(k0, k1, m-1, 0, 0), W(k0, k1, m-1, 0, 1))
    _advec4_generic(g, ut, u, vx, vy, vz, g.dzi4, g.kstart)


def advec_4_v(g, vt, u, v, w):
    """src/advec_4.cxx:188-286"""
    TF = g.TF
    I = lambda a, b, c, d: _w4(CI, a, b, c, d, TF)
    U = lambda k0, k1, dk=0, dj=0, di=0: _S(g, u, dk, dj, di, k0, k1)
    V = lambda k0, k1, dk=0, dj=0, di=0: _S(g, v, dk, dj, di, k0, k1)
    W = lambda k0, k1, dk=0, dj=0, di=0: _S(g, w, dk, dj, di, k0, k1)
    vx = lambda m, k0, k1: I(U(k0, k1, 0, -2, m-1), U(k0, k1, 0, -1, m-1), U(k0, k1, 0, 0, m-1), U(k0, k1, 0, 1, m-1))
    vy = lambda m, k0, k1: I(V(k0, k1, 0, m-3), V(k0, k1, 0, m-2), V(k0, k1, 0, m-1), V(k0, k1, 0, m))
    vz = lambda m, k0, k1: I(W(k0, k1, m-1, -2), W(k0, k1, m-1, -1), W(k0, k1, m-1, 0), W(k0, k1, m-1, 1))
    _advec4_generic(g, vt, v, vx, vy, vz, g.dzi4, g.kstart)


def advec_4_w(g, wt, u, v, w):
    """src/advec_4.cxx:288-386 (rows kstart+1 .. kend-1; both factors of the outermost vertical flux one-sided)"""
    TF = g.TF
    I = lambda a, b, c, d: _w4(CI, a, b, c, d, TF)
    U = lambda k0, k1, dk=0, dj=0, di=0: _S(g, u, dk, dj, di, k0, k1)
    V = lambda k0, k1, dk=0, dj=0, di=0: _S(g, v, dk, dj, di, k0, k1)
    W = lambda k0, k1, dk=0, dj=0, di=0: _S(g, w, dk, dj, di, k0, k1)
    vx = lambda m, k0, k1: I(U(k0, k1, -2, 0, m-1), U(k0, k1, -1, 0, m-1), U(k0, k1, 0, 0, m-1), U(k0, k1, 1, 0, m-1))
    vy = lambda m, k0, k1: I(V(k0, k1, -2, m-1), V(k0, k1, -1, m-1), V(k0, k1, 0, m-1), V(k0, k1, 1, m-1))
    lo = g.kstart+1
    def vz(m, k0, k1):
        if k0 == lo and m == 0:
            return _w4(BI, W(k0, k1, -2), W(k0, k1, -1), W(k0, k1, 0), W(k0, k1, 1), TF)
        if k1 == g.kend and m == 3:
            return _w4(TI, W(k0, k1, -1), W(k0, k1, 0), W(k0, k1, 1), W(k0, k1, 2), TF)
        return I(W(k0, k1, m-3), W(k0, k1, m-2), W(k0, k1, m-1), W(k0, k1, m))
    _advec4_generic(g, wt, w, vx, vy, vz, g.dzhi4, lo)


def advec_4_s(g, st, s, u, v, w):
    """src/advec_4.cxx:388-486 (face velocities used directly)"""
    vx = lambda m, k0, k1: _S(g, u, 0, 0, m-1, k0, k1)
    vy = lambda m, k0, k1: _S(g, v, 0, m-1, 0, k0, k1)
    vz = lambda m, k0, k1: _S(g, w, m-1, 0, 0, k0, k1)
    _advec4_generic(g, st, s, vx, vy, vz, g.dzi4, g.kstart)


# --------------------------------------------------------------------------------------
# Advec_4m (reference src/advec_4m.cxx:88-480): fully conservative 4th-order advection.  Every direction is
#   - grad4( V(-1) * interp2(q[-3], q[0]),  V(0) * interp2(q[-1], q[0]),  V(+1) * interp2(q[0], q[+1]),  V(+2) * interp2(q[0], q[+3]) )
# with V the advecting velocity interpolated (interp4c) to the four flux points; the whole right-hand side is ONE `+=` statement.
# At the first / last row the outermost vertical term is mirrored over the wall (no-penetration).
# --------------------------------------------------------------------------------------
def _i4c(TF, a, b, c, d):
    """interp4c (include/finite_difference.h:94-97): ci0*(a+d) + ci1*(b+c)"""
    return TF(CI[0])*(a + d) + TF(CI[1])*(b + c)


def _g4(TF, a, b, c, d):
    """grad4 (include/finite_difference.h:128-131): -cg0*(d-a) - cg1*(c-b)"""
    return -TF(CG[0])*(d - a) - TF(CG[1])*(c - b)


def _advec4m_generic(g, at, q, velx, vely, velz, dzx, lo, wall_rows=True):
    TF = g.TF
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    h = TF(0.5)
    i2 = lambda a, b: h*(a + b)
    rows = _advec4_rows(g, lo) if wall_rows else ((lo, g.kend, False, False),)
    for (k0, k1, bot, top) in rows:
        if k1 <= k0:
            continue
        Q = lambda dk=0, dj=0, di=0: _S(g, q, dk, dj, di, k0, k1)
        tgt = _S(g, at, 0, 0, 0, k0, k1)
        fx = _g4(TF, velx(0, k0, k1)*i2(Q(0, 0, -3), Q()), velx(1, k0, k1)*i2(Q(0, 0, -1), Q()),
                     velx(2, k0, k1)*i2(Q(), Q(0, 0, 1)), velx(3, k0, k1)*i2(Q(), Q(0, 0, 3)))
        fy = _g4(TF, vely(0, k0, k1)*i2(Q(0, -3), Q()), vely(1, k0, k1)*i2(Q(0, -1), Q()),
                     vely(2, k0, k1)*i2(Q(), Q(0, 1)), vely(3, k0, k1)*i2(Q(), Q(0, 3)))
        if bot:
            z0 = -velz(2, k0, k1)*i2(Q(-1), Q(2))
        else:
            z0 = velz(0, k0, k1)*i2(Q(-3), Q())
        if top:
            z3 = -velz(1, k0, k1)*i2(Q(-2), Q(1))
        else:
            z3 = velz(3, k0, k1)*i2(Q(), Q(3))
        fz = _g4(TF, z0, velz(1, k0, k1)*i2(Q(-1), Q()), velz(2, k0, k1)*i2(Q(), Q(1)), z3)
        tgt[...] += - fx*dxi - fy*dyi - fz*_K(g, dzx, 0, k0, k1)


def advec_4m_u(g, ut, u, v, w):
    """src/advec_4m.cxx:90-182"""
    TF = g.TF
    U = lambda k0, k1, dk=0, dj=0, di=0: _S(g, u, dk, dj, di, k0, k1)
    V = lambda k0, k1, dk=0, dj=0, di=0: _S(g, v, dk, dj, di, k0, k1)
    W = lambda k0, k1, dk=0, dj=0, di=0: _S(g, w, dk, dj, di, k0, k1)
    vx = lambda m, k0, k1: _i4c(TF, U(k0, k1, 0, 0, m-3), U(k0, k1, 0, 0, m-2), U(k0, k1, 0, 0, m-1), U(k0, k1, 0, 0, m))
    vy = lambda m, k0, k1: _i4c(TF, V(k0, k1, 0, m-1, -2), V(k0, k1, 0, m-1, -1), V(k0, k1, 0, m-1, 0), V(k0, k1, 0, m-1, 1))
    vz = lambda m, k0, k1: _i4c(TF, W(k0, k1, m-1, 0, -2), W(k0, k1, m-1, 0, -1), W(k0, k1, m-1, 0, 0), W(k0, k1, m-1, 0, 1))
    _advec4m_generic(g, ut, u, vx, vy, vz, g.dzi4, g.kstart)


def advec_4m_v(g, vt, u, v, w):
    """src/advec_4m.cxx:184-276"""
    TF = g.TF
    U = lambda k0, k1, dk=0, dj=0, di=0: _S(g, u, dk, dj, di, k0, k1)
    V = lambda k0, k1, dk=0, dj=0, di=0: _S(g, v, dk, dj, di, k0, k1)
    W = lambda k0, k1, dk=0, dj=0, di=0: _S(g, w, dk, dj, di, k0, k1)
    vx = lambda m, k0, k1: _i4c(TF, U(k0, k1, 0, -2, m-1), U(k0, k1, 0, -1, m-1), U(k0, k1, 0, 0, m-1), U(k0, k1, 0, 1, m-1))
    vy = lambda m, k0, k1: _i4c(TF, V(k0, k1, 0, m-3), V(k0, k1, 0, m-2), V(k0, k1, 0, m-1), V(k0, k1, 0, m))
    vz = lambda m, k0, k1: _i4c(TF, W(k0, k1, m-1, -2), W(k0, k1, m-1, -1), W(k0, k1, m-1, 0), W(k0, k1, m-1, 1))
    _advec4m_generic(g, vt, v, vx, vy, vz, g.dzi4, g.kstart)


def advec_4m_w(g, wt, u, v, w):
    """src/advec_4m.cxx:278-323 (rows kstart+1 .. kend-1, no wall rows: w at the walls is zero)"""
    TF = g.TF
    U = lambda k0, k1, dk=0, dj=0, di=0: _S(g, u, dk, dj, di, k0, k1)
    V = lambda k0, k1, dk=0, dj=0, di=0: _S(g, v, dk, dj, di, k0, k1)
    W = lambda k0, k1, dk=0, dj=0, di=0: _S(g, w, dk, dj, di, k0, k1)
    vx = lambda m, k0, k1: _i4c(TF, U(k0, k1, -2, 0, m-1), U(k0, k1, -1, 0, m-1), U(k0, k1, 0, 0, m-1), U(k0, k1, 1, 0, m-1))
    vy = lambda m, k0, k1: _i4c(TF, V(k0, k1, -2, m-1), V(k0, k1, -1, m-1), V(k0, k1, 0, m-1), V(k0, k1, 1, m-1))
    vz = lambda m, k0, k1: _i4c(TF, W(k0, k1, m-3), W(k0, k1, m-2), W(k0, k1, m-1), W(k0, k1, m))
    _advec4m_generic(g, wt, w, vx, vy, vz, g.dzhi4, g.kstart+1, wall_rows=False)


def advec_4m_s(g, st, s, u, v, w):
    """src/advec_4m.cxx:325-400 (face velocities used directly)"""
    vx = lambda m, k0, k1: _S(g, u, 0, 0, m-1, k0, k1)
    vy = lambda m, k0, k1: _S(g, v, 0, m-1, 0, k0, k1)
    vz = lambda m, k0, k1: _S(g, w, m-1, 0, 0, k0, k1)
    _advec4m_generic(g, st, s, vx, vy, vz, g.dzi4, g.kstart)


def advec_4_cfl(g, u, v, w, dt):
    """src/advec_4.cxx:50-86: interp4c(a,b,c,d) = ci0*(a+d) + ci1*(b+c)"""
    TF = g.TF
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    i4c = lambda a, b, c, d: TF(CI[0])*(a+d) + TF(CI[1])*(b+c)
    c = ( np.abs(i4c(_S(g, u, 0, 0, -1), _S(g, u), _S(g, u, 0, 0, 1), _S(g, u, 0, 0, 2)))*dxi
        + np.abs(i4c(_S(g, v, 0, -1), _S(g, v), _S(g, v, 0, 1), _S(g, v, 0, 2)))*dyi
        + np.abs(i4c(_S(g, w, -1), _S(g, w), _S(g, w, 1), _S(g, w, 2)))*_K(g, g.dzi) )
    return TF(TF(c.max())*TF(dt))


def advec_4m_cfl(g, u, v, w, dt):
    """src/advec_4m.cxx:51-88: the interpolation is the plain four-term sum ci0*a + ci1*b + ci2*c + ci3*d (not interp4c)"""
    TF = g.TF
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    c0, c1, c2, c3 = (TF(x) for x in CI)
    s4 = lambda a, b, c, d: c0*a + c1*b + c2*c + c3*d
    c = ( np.abs(s4(_S(g, u, 0, 0, -1), _S(g, u), _S(g, u, 0, 0, 1), _S(g, u, 0, 0, 2)))*dxi
        + np.abs(s4(_S(g, v, 0, -1), _S(g, v), _S(g, v, 0, 1), _S(g, v, 0, 2)))*dyi
        + np.abs(s4(_S(g, w, -1), _S(g, w), _S(g, w, 1), _S(g, w, 2)))*_K(g, g.dzi) )
    return TF(TF(c.max())*TF(dt))


# --------------------------------------------------------------------------------------
# Diff_4 (reference src/diff_4.cxx:40-175): nu * (7-point cdg laplacian in x, y; div(grad) with cg weights and
# one-sided bg / tg gradients at the walls in z).  Three separate `+=` statements per point.
# --------------------------------------------------------------------------------------
def _diff4(g, at, a, visc, lo, dz_in, dz_out, dxidxi, dyidyi, in_off):
    TF = g.TF
    visc = TF(visc)
    dim3 = g.jtot > 1
    cdg = [TF(c) for c in CDG]
    for (k0, k1, bot, top) in _advec4_rows(g, lo):
        if k1 <= k0:
            continue
        A = lambda dk=0, dj=0, di=0: _S(g, a, dk, dj, di, k0, k1)
        tgt = _S(g, at, 0, 0, 0, k0, k1)
        tgt[...] += visc * (cdg[3]*A(0, 0, -3) + cdg[2]*A(0, 0, -2) + cdg[1]*A(0, 0, -1) + cdg[0]*A()
                            + cdg[1]*A(0, 0, 1) + cdg[2]*A(0, 0, 2) + cdg[3]*A(0, 0, 3)) * dxidxi
        if dim3:
            tgt[...] += visc * (cdg[3]*A(0, -3) + cdg[2]*A(0, -2) + cdg[1]*A(0, -1) + cdg[0]*A()
                                + cdg[1]*A(0, 1) + cdg[2]*A(0, 2) + cdg[3]*A(0, 3)) * dyidyi
        def grad(m):
            if bot and m == 0:
                return _w4(BG, A(-2), A(-1), A(0), A(1), TF)
            if top and m == 3:
                return _w4(TG, A(-1), A(0), A(1), A(2), TF)
            return _w4(CG, A(m-3), A(m-2), A(m-1), A(m), TF)
        Kin = lambda m: _K(g, dz_in, m - 1 + in_off, k0, k1)
        tgt[...] += visc * ( TF(CG[0])*grad(0)*Kin(0) + TF(CG[1])*grad(1)*Kin(1)
                           + TF(CG[2])*grad(2)*Kin(2) + TF(CG[3])*grad(3)*Kin(3) ) * _K(g, dz_out, 0, k0, k1)


def diff_4_c(g, at, a, visc):
    """src/diff_4.cxx:40-105 (`dxidxi = 1./(dx*dx)`: double division narrowed to TF)"""
    TF = g.TF
    _diff4(g, at, a, visc, g.kstart, g.dzhi4, g.dzi4,
           TF(1./np.float64(g.dx*g.dx)), TF(1./np.float64(g.dy*g.dy)), 0)


def diff_4_w(g, wt, w, visc):
    """src/diff_4.cxx:107-175 (`dxidxi = 1/(dx*dx)`: TF division; inner metric dzi4[k-2..k+1], outer dzhi4[k])"""
    TF = g.TF
    _diff4(g, wt, w, visc, g.kstart+1, g.dzi4, g.dzhi4,
           TF(1.)/TF(g.dx*g.dx), TF(1.)/TF(g.dy*g.dy), -1)


# --------------------------------------------------------------------------------------
# Diff_smag2 (reference include/diff_kernels.h:34-511, src/diff_smag2.cxx:148-269)
# --------------------------------------------------------------------------------------
def _load_cport():
    import ctypes, os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liboracle_c.so")
    return ctypes.CDLL(p) if os.path.exists(p) else None

_CPORT = _load_cport()

def _libm_pow(x, y):
    """std::pow(TF, TF) through the C library (oracle/cport/libm_vec.c), the libm the reference's CPU
    build links; numpy's own SIMD pow is 1 ulp apart now and then.  Falls back to numpy when the
    helper has not been built (make -C oracle cport): parity to ~1 ulp instead of bit-exact."""
    import ctypes
    x = np.ascontiguousarray(x)
    if _CPORT is None or x.dtype not in (np.float64, np.float32):
        return np.power(x, x.dtype.type(y))
    out = np.empty_like(x)
    if x.dtype == np.float64:
        _CPORT.vpow_f64(x.ctypes.data_as(ctypes.c_void_p), ctypes.c_double(float(y)), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(x.size))
    else:
        _CPORT.vpow_f32(x.ctypes.data_as(ctypes.c_void_p), ctypes.c_float(float(y)), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(x.size))
    return out

DSMALL = 1.e-9   # Constants::dsmall (reference include/constants.h)
KAPPA = 0.4      # Constants::kappa
GRAV = 9.81      # Constants::grav

def _libm_exp(x):
    """std::exp(TF) through the C library (see _libm_pow)"""
    import ctypes
    x = np.ascontiguousarray(x)
    if _CPORT is None or not hasattr(_CPORT, "vexp_f64") or x.dtype not in (np.float64, np.float32):
        return np.exp(x)
    out = np.empty_like(x)
    fn = _CPORT.vexp_f64 if x.dtype == np.float64 else _CPORT.vexp_f32
    fn(x.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(x.size))
    return out


def _pow2(a):
    return a*a

def _add_dsmall(val):
    # `strain2[ijk] += Constants::dsmall` with a *double* constant: the sum is formed in double
    # and rounded back to TF (include/diff_kernels.h:101,140; include/constants.h:97).
    return (val.astype(np.float64) + DSMALL).astype(val.dtype)

def diff_strain2(g, strain2, u, v, w, ugradbot, vgradbot, surface):
    """include/diff_kernels.h:34-142"""
    TF = g.TF
    ks, ke = g.kstart, g.kend
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    e = TF(0.125); h = TF(0.5); two = TF(2.)
    k_offset = 1 if surface else 0

    def common(k0, k1):
        U = lambda di=0, dj=0, dk=0: _S(g, u, dk, dj, di, k0, k1)
        V = lambda di=0, dj=0, dk=0: _S(g, v, dk, dj, di, k0, k1)
        W = lambda di=0, dj=0, dk=0: _S(g, w, dk, dj, di, k0, k1)
        return U, V, W

    if surface:
        k0, k1 = ks, ks+1
        U, V, W = common(k0, k1)
        ij = (slice(g.jstart, g.jend), slice(g.istart, g.iend))
        val = two*(
            + _pow2((U(1)-U(0))*dxi)
            + _pow2((V(0,1)-V(0,0))*dyi)
            + _pow2((W(0,0,1)-W(0,0,0))*g.dzi[ks])
            + e*_pow2((U(0,0)-U(0,-1))*dyi + (V(0,0)-V(-1,0))*dxi)
            + e*_pow2((U(1,0)-U(1,-1))*dyi + (V(1,0)-V(0,0))*dxi)
            + e*_pow2((U(0,1)-U(0,0))*dyi + (V(0,1)-V(-1,1))*dxi)
            + e*_pow2((U(1,1)-U(1,0))*dyi + (V(1,1)-V(0,1))*dxi)
            + h*_pow2(ugradbot[ij])
            + e*_pow2((W(0,0,0)-W(-1,0,0))*dxi)
            + e*_pow2((W(1,0,0)-W(0,0,0))*dxi)
            + e*_pow2((W(0,0,1)-W(-1,0,1))*dxi)
            + e*_pow2((W(1,0,1)-W(0,0,1))*dxi)
            + h*_pow2(vgradbot[ij])
            + e*_pow2((W(0,0,0)-W(0,-1,0))*dyi)
            + e*_pow2((W(0,1,0)-W(0,0,0))*dyi)
            + e*_pow2((W(0,0,1)-W(0,-1,1))*dyi)
            + e*_pow2((W(0,1,1)-W(0,0,1))*dyi) )
        _S(g, strain2, 0, 0, 0, k0, k1)[...] = _add_dsmall(val)

    k0, k1 = ks + k_offset, ke
    U, V, W = common(k0, k1)
    dzi = _K(g, g.dzi, 0, k0, k1)
    dzhi0 = _K(g, g.dzhi, 0, k0, k1)
    dzhi1 = _K(g, g.dzhi, 1, k0, k1)
    val = two*(
        + _pow2((U(1)-U(0))*dxi)
        + _pow2((V(0,1)-V(0,0))*dyi)
        + _pow2((W(0,0,1)-W(0,0,0))*dzi)
        + e*_pow2((U(0,0)-U(0,-1))*dyi + (V(0,0)-V(-1,0))*dxi)
        + e*_pow2((U(1,0)-U(1,-1))*dyi + (V(1,0)-V(0,0))*dxi)
        + e*_pow2((U(0,1)-U(0,0))*dyi + (V(0,1)-V(-1,1))*dxi)
        + e*_pow2((U(1,1)-U(1,0))*dyi + (V(1,1)-V(0,1))*dxi)
        + e*_pow2((U(0,0,0)-U(0,0,-1))*dzhi0 + (W(0,0,0)-W(-1,0,0))*dxi)
        + e*_pow2((U(1,0,0)-U(1,0,-1))*dzhi0 + (W(1,0,0)-W(0,0,0))*dxi)
        + e*_pow2((U(0,0,1)-U(0,0,0))*dzhi1 + (W(0,0,1)-W(-1,0,1))*dxi)
        + e*_pow2((U(1,0,1)-U(1,0,0))*dzhi1 + (W(1,0,1)-W(0,0,1))*dxi)
        + e*_pow2((V(0,0,0)-V(0,0,-1))*dzhi0 + (W(0,0,0)-W(0,-1,0))*dyi)
        + e*_pow2((V(0,1,0)-V(0,1,-1))*dzhi0 + (W(0,1,0)-W(0,0,0))*dyi)
        + e*_pow2((V(0,0,1)-V(0,0,0))*dzhi1 + (W(0,0,1)-W(0,-1,1))*dyi)
        + e*_pow2((V(0,1,1)-V(0,1,0))*dzhi1 + (W(0,1,1)-W(0,0,1))*dyi) )
    _S(g, strain2, 0, 0, 0, k0, k1)[...] = _add_dsmall(val)


def diff_evisc(g, evisc, N2, bgradbot, z0m, cs, tPr, surface, mason=True):
    """src/diff_smag2.cxx:148-269 (calc_evisc; ends with boundary_cyclic.exec(evisc))"""
    TF = g.TF
    ks, ke = g.kstart, g.kend
    cs, tPr = TF(cs), TF(tPr)
    one_m = TF(1. - DSMALL)
    third = TF(1./3.)
    mlen0_k = (cs*_libm_pow((g.dx*g.dy*g.dz).astype(TF), third)).astype(TF)
    ij = (slice(g.jstart, g.jend), slice(g.istart, g.iend))
    if not surface:
        k0, k1 = ks, ke
        fac = _K(g, _pow2(mlen0_k), 0, k0, k1)
        ev = _S(g, evisc, 0, 0, 0, k0, k1)
        rit = _S(g, N2, 0, 0, 0, k0, k1) / ev / tPr
        rit = np.minimum(rit, one_m)
        ev[...] = fac * np.sqrt(ev) * np.sqrt(TF(1.) - rit)
        evisc[ks-1] = evisc[ks]
        evisc[ke] = evisc[ke-1]
    else:
        n_mason = TF(2.)
        def mlen_of(k0, k1):
            m0 = _K(g, mlen0_k, 0, k0, k1)
            if not mason:
                return m0 + np.zeros_like(_S(g, evisc, 0, 0, 0, k0, k1))
            zz = _K(g, g.z, 0, k0, k1) + z0m[ij][None, :, :]
            return _libm_pow(TF(1.)/(TF(1.)/_libm_pow(m0 + np.zeros_like(zz), n_mason) + TF(1.)/(_libm_pow(TF(KAPPA)*zz, n_mason))), TF(1.)/n_mason)
        # bottom (:216-239)
        k0, k1 = ks, ks+1
        ev = _S(g, evisc, 0, 0, 0, k0, k1)
        rit = bgradbot[ij][None, :, :] / ev / tPr
        rit = np.minimum(rit, one_m)
        ev[...] = _pow2(mlen_of(k0, k1)) * np.sqrt(ev) * np.sqrt(TF(1.) - rit)
        # interior (:241-265)
        k0, k1 = ks+1, ke
        ev = _S(g, evisc, 0, 0, 0, k0, k1)
        rit = _S(g, N2, 0, 0, 0, k0, k1) / ev / tPr
        rit = np.minimum(rit, one_m)
        ev[...] = _pow2(mlen_of(k0, k1)) * np.sqrt(ev) * np.sqrt(TF(1.) - rit)
    boundary_cyclic(g, evisc)


def diff_evisc_neutral(g, evisc, u, v, w, z0m, cs, visc, surface, mason=True):
    """src/diff_smag2.cxx:47-146: no stability correction (swthermo=0).  `evisc` holds strain2 on entry.  Surface model:
    Mason wall correction with n = 1; resolved walls: van Driest damping from the wall shear of either wall, then the
    mirror over the walls; ends with the cyclic fill."""
    TF = g.TF
    ks, ke = g.kstart, g.kend
    third = TF(1./3.)
    mlen0_k = (TF(cs)*_libm_pow((g.dx*g.dy*g.dz).astype(TF), third)).astype(TF)
    ij = (slice(g.jstart, g.jend), slice(g.istart, g.iend))
    ev = _S(g, evisc)
    if not surface:
        visc = TF(visc); A = TF(26.)
        def u_tau(kw):
            du = visc*(u[kw][ij] - u[kw-1][ij])*g.dzhi[kw]
            dv = visc*(v[kw][ij] - v[kw-1][ij])*g.dzhi[kw]
            return _libm_pow((_pow2(du) + _pow2(dv)).astype(TF), TF(0.25))
        ut_bot, ut_top = u_tau(ks)[None, :, :], u_tau(ke)[None, :, :]
        zk = _K(g, g.z)
        fac_bot = TF(1.) - _libm_exp((-(zk*ut_bot) / (A*visc)).astype(TF))
        fac_top = TF(1.) - _libm_exp((-((g.zsize - zk)*ut_top) / (A*visc)).astype(TF))
        fac = np.minimum(fac_bot, fac_top)
        ev[...] = _pow2(fac * _K(g, mlen0_k)) * np.sqrt(ev)
        evisc[ks-1] = evisc[ks]
        evisc[ke] = evisc[ke-1]
    else:
        if mason:
            one = TF(1.)      # n_mason = 1: pow(x, 1) is exact, kept for the reference's expression structure
            zz = (_K(g, g.z) + z0m[ij][None, :, :]).astype(TF)
            m0 = _K(g, mlen0_k) + np.zeros_like(zz)
            mlen = _libm_pow(one/(one/_libm_pow(m0, one) + one/(_libm_pow(TF(KAPPA)*zz, one))), one/one)
        else:
            mlen = _K(g, mlen0_k)
        ev[...] = _pow2(mlen) * np.sqrt(ev)
    boundary_cyclic(g, evisc)


def _diff_uv(g, at, a, b, w, evisc, fluxbot, fluxtop, rhoref, rhorefh, visc, surface, is_u):
    """diff_u (include/diff_kernels.h:144-244) / diff_v (:246-347); `a` is the component
    being diffused, `b` the other horizontal component.  For diff_v the roles of x and y swap."""
    TF = g.TF
    ks, ke = g.kstart, g.kend
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    visc = TF(visc)
    q = TF(0.25); two = TF(2.)
    ij = (slice(g.jstart, g.jend), slice(g.istart, g.iend))

    def sh(arr, k0, k1):
        # offsets given as (d_along, d_cross, dk): along = own direction of component a
        if is_u:
            return lambda da=0, dc=0, dk=0: _S(g, arr, dk, dc, da, k0, k1)
        return lambda da=0, dc=0, dk=0: _S(g, arr, dk, da, dc, k0, k1)
    d_al, d_cr = (dxi, dyi) if is_u else (dyi, dxi)

    def horizontal(k0, k1):
        E = sh(evisc, k0, k1); A = sh(a, k0, k1); B = sh(b, k0, k1)
        if is_u:
            ev_p = E(0) + visc                       # evisce
            ev_m = E(-1) + visc                      # eviscw
            ev_cp = q*(E(-1,0) + E(0,0) + E(-1,1) + E(0,1)) + visc     # eviscn
            ev_cm = q*(E(-1,-1) + E(0,-1) + E(-1,0) + E(0,0)) + visc   # eviscs
            return (
                + ( ev_p*(A(1)-A(0))*d_al - ev_m*(A(0)-A(-1))*d_al ) * two*d_al
                + ( ev_cp*((A(0,1)-A(0,0))*d_cr + (B(0,1)-B(-1,1))*d_al)
                  - ev_cm*((A(0,0)-A(0,-1))*d_cr + (B(0,0)-B(-1,0))*d_al) ) * d_cr )
        else:
            # diff_v: evisce/eviscw are the 4-point means (cross = x), eviscn/s plain
            # reference order: evisc[ijk-jj] + evisc[ijk] + evisc[ijk+ii-jj] + evisc[ijk+ii]
            ev_e = q*(E(-1,0) + E(0,0) + E(-1,1) + E(0,1)) + visc
            ev_w = q*(E(-1,-1) + E(0,-1) + E(-1,0) + E(0,0)) + visc
            ev_n = E(0) + visc
            ev_s = E(-1) + visc
            return (
                + ( ev_e*((A(0,1)-A(0,0))*d_cr + (B(0,1)-B(-1,1))*d_al)
                  - ev_w*((A(0,0)-A(0,-1))*d_cr + (B(0,0)-B(-1,0))*d_al) ) * d_cr
                + ( ev_n*(A(1)-A(0))*d_al - ev_s*(A(0)-A(-1))*d_al ) * two*d_al )

    def ev_t(k0, k1):
        E = sh(evisc, k0, k1)
        return q*(E(-1,0,0) + E(0,0,0) + E(-1,0,1) + E(0,0,1)) + visc
    def ev_b(k0, k1):
        E = sh(evisc, k0, k1)
        return q*(E(-1,0,-1) + E(0,0,-1) + E(-1,0,0) + E(0,0,0)) + visc
    def top_flux(k0, k1):
        A = sh(a, k0, k1); W = sh(w, k0, k1)
        return _K(g, rhorefh, 1, k0, k1) * ev_t(k0, k1)*((A(0,0,1)-A(0,0,0))*_K(g, g.dzhi, 1, k0, k1) + (W(0,0,1)-W(-1,0,1))*d_al)
    def bot_flux(k0, k1):
        A = sh(a, k0, k1); W = sh(w, k0, k1)
        return _K(g, rhorefh, 0, k0, k1) * ev_b(k0, k1)*((A(0,0,0)-A(0,0,-1))*_K(g, g.dzhi, 0, k0, k1) + (W(0,0,0)-W(-1,0,0))*d_al)

    k_offset = 1 if surface else 0
    if surface:
        k0, k1 = ks, ks+1
        _S(g, at, 0, 0, 0, k0, k1)[...] += (
            horizontal(k0, k1)
            + ( top_flux(k0, k1) + rhorefh[ks] * fluxbot[ij][None] ) / rhoref[ks] * g.dzi[ks] )
        k0, k1 = ke-1, ke
        _S(g, at, 0, 0, 0, k0, k1)[...] += (
            horizontal(k0, k1)
            + ( - rhorefh[ke] * fluxtop[ij][None] - bot_flux(k0, k1) ) / rhoref[ke-1] * g.dzi[ke-1] )
    k0, k1 = ks + k_offset, ke - k_offset
    if k1 > k0:
        _S(g, at, 0, 0, 0, k0, k1)[...] += (
            horizontal(k0, k1)
            + ( top_flux(k0, k1) - bot_flux(k0, k1) ) / _K(g, rhoref, 0, k0, k1) * _K(g, g.dzi, 0, k0, k1) )


def diff_u(g, ut, u, v, w, evisc, fluxbot, fluxtop, rhoref, rhorefh, visc, surface):
    _diff_uv(g, ut, u, v, w, evisc, fluxbot, fluxtop, rhoref, rhorefh, visc, surface, True)

def diff_v(g, vt, u, v, w, evisc, fluxbot, fluxtop, rhoref, rhorefh, visc, surface):
    _diff_uv(g, vt, v, u, w, evisc, fluxbot, fluxtop, rhoref, rhorefh, visc, surface, False)


def diff_w(g, wt, u, v, w, evisc, rhoref, rhorefh, visc):
    """include/diff_kernels.h:349-392"""
    TF = g.TF
    ks, ke = g.kstart, g.kend
    dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
    visc = TF(visc); q = TF(0.25); two = TF(2.)
    k0, k1 = ks+1, ke
    E = lambda di=0, dj=0, dk=0: _S(g, evisc, dk, dj, di, k0, k1)
    U = lambda di=0, dj=0, dk=0: _S(g, u, dk, dj, di, k0, k1)
    V = lambda di=0, dj=0, dk=0: _S(g, v, dk, dj, di, k0, k1)
    W = lambda di=0, dj=0, dk=0: _S(g, w, dk, dj, di, k0, k1)
    evisce = q*(E(0,0,-1) + E(0,0,0) + E(1,0,-1) + E(1,0,0)) + visc
    eviscw = q*(E(-1,0,-1) + E(-1,0,0) + E(0,0,-1) + E(0,0,0)) + visc
    eviscn = q*(E(0,0,-1) + E(0,0,0) + E(0,1,-1) + E(0,1,0)) + visc
    eviscs = q*(E(0,-1,-1) + E(0,-1,0) + E(0,0,-1) + E(0,0,0)) + visc
    evisct = E(0,0,0) + visc
    eviscb = E(0,0,-1) + visc
    dzhi = _K(g, g.dzhi, 0, k0, k1)
    _S(g, wt, 0, 0, 0, k0, k1)[...] += (
        + ( evisce*((W(1)-W(0))*dxi + (U(1,0,0)-U(1,0,-1))*dzhi)
          - eviscw*((W(0)-W(-1))*dxi + (U(0,0,0)-U(0,0,-1))*dzhi) ) * dxi
        + ( eviscn*((W(0,1)-W(0,0))*dyi + (V(0,1,0)-V(0,1,-1))*dzhi)
          - eviscs*((W(0,0)-W(0,-1))*dyi + (V(0,0,0)-V(0,0,-1))*dzhi) ) * dyi
        + ( _K(g, rhoref, 0, k0, k1) * evisct*(W(0,0,1)-W(0,0,0))*_K(g, g.dzi, 0, k0, k1)
          - _K(g, rhoref,-1, k0, k1) * eviscb*(W(0,0,0)-W(0,0,-1))*_K(g, g.dzi,-1, k0, k1) ) / _K(g, rhorefh, 0, k0, k1) * two*dzhi )


def diff_c(g, at, a, evisc, fluxbot, fluxtop, rhoref, rhorefh, tPr, visc, surface):
    """include/diff_kernels.h:394-484.  dxidxi = 1./(dx*dx) computed in double then cast (src/diff_smag2.cxx:445)."""
    TF = g.TF
    ks, ke = g.kstart, g.kend
    dxidxi = TF(1./(float(g.dx)*float(g.dx))); dyidyi = TF(1./(float(g.dy)*float(g.dy)))
    visc = TF(visc); h = TF(0.5)
    tPr_i = TF(1)/TF(tPr)
    ij = (slice(g.jstart, g.jend), slice(g.istart, g.iend))

    def parts(k0, k1):
        E = lambda di=0, dj=0, dk=0: _S(g, evisc, dk, dj, di, k0, k1)
        A = lambda di=0, dj=0, dk=0: _S(g, a, dk, dj, di, k0, k1)
        evisce = h*(E(0)+E(1)) * tPr_i + visc
        eviscw = h*(E(-1)+E(0)) * tPr_i + visc
        eviscn = h*(E(0,0)+E(0,1)) * tPr_i + visc
        eviscs = h*(E(0,-1)+E(0,0)) * tPr_i + visc
        hor_x = ( evisce*(A(1)-A(0)) - eviscw*(A(0)-A(-1)) ) * dxidxi
        hor_y = ( eviscn*(A(0,1)-A(0,0)) - eviscs*(A(0,0)-A(0,-1)) ) * dyidyi
        top = lambda: _K(g, rhorefh, 1, k0, k1) * (h*(E(0,0,0)+E(0,0,1)) * tPr_i + visc)*(A(0,0,1)-A(0,0,0))*_K(g, g.dzhi, 1, k0, k1)
        bot = lambda: _K(g, rhorefh, 0, k0, k1) * (h*(E(0,0,-1)+E(0,0,0)) * tPr_i + visc)*(A(0,0,0)-A(0,0,-1))*_K(g, g.dzhi, 0, k0, k1)
        return hor_x, hor_y, top, bot

    k_offset = 1 if surface else 0
    if surface:
        k0, k1 = ks, ks+1
        hx, hy, top, bot = parts(k0, k1)
        _S(g, at, 0, 0, 0, k0, k1)[...] += (
            + hx + hy + ( top() + rhorefh[ks] * fluxbot[ij][None] ) / rhoref[ks] * g.dzi[ks] )
        k0, k1 = ke-1, ke
        hx, hy, top, bot = parts(k0, k1)
        _S(g, at, 0, 0, 0, k0, k1)[...] += (
            + hx + hy + ( -rhorefh[ke] * fluxtop[ij][None] - bot() ) / rhoref[ke-1] * g.dzi[ke-1] )
    k0, k1 = ks + k_offset, ke - k_offset
    if k1 > k0:
        hx, hy, top, bot = parts(k0, k1)
        _S(g, at, 0, 0, 0, k0, k1)[...] += (
            + hx + hy + ( top() - bot() ) / _K(g, rhoref, 0, k0, k1) * _K(g, g.dzi, 0, k0, k1) )


def diff_dnmul(g, evisc, tPr):
    """include/diff_kernels.h:486-511; dxidxi passed as double 1./(dx*dx) cast to TF (src/diff_smag2.cxx:317-326)"""
    TF = g.TF
    dxidxi = TF(1./(float(g.dx)*float(g.dx))); dyidyi = TF(1./(float(g.dy)*float(g.dy)))
    tPrfac_i = TF(1)/min(TF(1.), TF(tPr))
    dzi = _K(g, g.dzi)
    return TF(np.abs(_S(g, evisc)*tPrfac_i*(dxidxi + dyidyi + dzi*dzi)).max())


# --------------------------------------------------------------------------------------
# Diff_tke2: Deardorff (1980) SGS-TKE closure (reference src/diff_tke2.cxx:48-512) and the Limiter's
# tendency_limiter (src/limiter.cxx:35-59).  Surface model only ("Resolved wall not supported").
# --------------------------------------------------------------------------------------
SGSTKE_MIN = 1.e-7    # Constants::sgstke_min (include/constants.h:59)

def _tke2_mlen0(g):
    """std::pow(dx*dy*dz[k], TF(1./3.)) per level (src/diff_tke2.cxx:104,167,188)"""
    TF = g.TF
    return _libm_pow((g.dx*g.dy*g.dz).astype(TF), TF(1./3.)).astype(TF)

def _tke2_ij(g):
    return (slice(g.jstart, g.jend), slice(g.istart, g.iend))

def tke2_enforce_min(g, sgstke):
    """src/diff_tke2.cxx:48-71"""
    TF = g.TF
    a = _S(g, sgstke)
    a[...] = np.maximum(a, TF(SGSTKE_MIN))
    boundary_cyclic(g, sgstke)

def tke2_evisc_neutral(g, evisc, sgstke, z0m, cn, cm, mason=True):
    """src/diff_tke2.cxx:73-134 (note: the Mason correction uses z[kstart] at every level, :118)"""
    TF = g.TF
    cm = TF(cm)
    m0 = _K(g, _tke2_mlen0(g))
    a = _S(g, sgstke)
    if mason:
        zz = TF(KAPPA)*(g.z[g.kstart] + z0m[_tke2_ij(g)])[None, :, :]
        fac = np.sqrt(TF(1.)/(TF(1.)/_pow2(m0) + TF(1.)/_pow2(zz)))
    else:
        fac = m0 + np.zeros_like(a)
    _S(g, evisc)[...] = cm*fac*np.sqrt(a)
    boundary_cyclic(g, evisc)

def _tke2_fac(g, a, N2, bgradbot, z0m, cn, mason, variant):
    """Length scale of calc_evisc / calc_evisc_heat (variant 0: cn*sqrt(a/N2), Mason by sqrt / pow2, src/diff_tke2.cxx:170-228)
    and of sgstke_diss_tend (variant 1: cn*sqrt(a)/sqrt(N2), Mason by std::pow, :420-466).  Returns (fac, mlen0) on the interior."""
    TF = g.TF
    cn = TF(cn)
    ks, ke = g.kstart, g.kend
    m0 = _K(g, _tke2_mlen0(g)) + np.zeros_like(_S(g, a))
    n2 = _S(g, N2).copy()
    n2[0] = bgradbot[_tke2_ij(g)]                   # lowest level: the surface model's db/dz
    aa = _S(g, a)
    with np.errstate(divide="ignore", invalid="ignore"):
        if variant == 0:
            mlen = cn*np.sqrt(aa/n2)
        else:
            mlen = cn*np.sqrt(aa)/np.sqrt(n2)
    mlen = np.where(n2 > 0, mlen, m0)
    fac = np.minimum(m0, mlen)
    if mason:
        zz = TF(KAPPA)*(_K(g, g.z) + z0m[_tke2_ij(g)][None, :, :])
        if variant == 0:
            fac = np.sqrt(TF(1.)/(TF(1.)/_pow2(fac) + TF(1.)/_pow2(zz)))
        else:
            two = TF(2.)
            fac = _libm_pow(TF(1.)/(TF(1.)/_libm_pow(fac, two) + TF(1.)/_libm_pow(zz, two)), TF(1.)/two)
    return fac, m0, n2

def tke2_evisc(g, evisc, sgstke, N2, bgradbot, z0m, cn, cm, mason=True):
    """src/diff_tke2.cxx:136-234"""
    TF = g.TF
    fac, _, _ = _tke2_fac(g, sgstke, N2, bgradbot, z0m, cn, mason, 0)
    _S(g, evisc)[...] = TF(cm)*fac*np.sqrt(_S(g, sgstke))
    boundary_cyclic(g, evisc)

def tke2_evisc_heat(g, evisch, evisc, sgstke, N2, bgradbot, z0m, cn, ch1, ch2, mason=True):
    """src/diff_tke2.cxx:236-332"""
    TF = g.TF
    fac, m0, _ = _tke2_fac(g, sgstke, N2, bgradbot, z0m, cn, mason, 0)
    _S(g, evisch)[...] = (TF(ch1) + TF(ch2)*fac/m0)*_S(g, evisc)
    boundary_cyclic(g, evisch)

def tke2_shear_tend(g, at, evisc, strain2):
    """src/diff_tke2.cxx:334-357"""
    _S(g, at)[...] += _S(g, evisc)*_S(g, strain2)

def tke2_buoy_tend(g, at, evisch, N2, bgradbot):
    """src/diff_tke2.cxx:359-393"""
    n2 = _S(g, N2).copy()
    n2[0] = bgradbot[_tke2_ij(g)]
    _S(g, at)[...] -= _S(g, evisch)*n2

def tke2_diss_tend(g, at, a, N2, bgradbot, z0m, cn, ce1, ce2, mason=True):
    """src/diff_tke2.cxx:395-469"""
    TF = g.TF
    fac, m0, _ = _tke2_fac(g, a, N2, bgradbot, z0m, cn, mason, 1)
    _S(g, at)[...] -= (TF(ce1) + TF(ce2)*fac/m0)*_libm_pow(_S(g, a), TF(3./2.))/fac

def tke2_diss_tend_neutral(g, at, a, z0m, ce1, ce2, mason=True):
    """src/diff_tke2.cxx:471-511"""
    TF = g.TF
    m0 = _K(g, _tke2_mlen0(g)) + np.zeros_like(_S(g, a))
    if mason:
        two = TF(2.)
        zz = TF(KAPPA)*(_K(g, g.z) + z0m[_tke2_ij(g)][None, :, :])
        fac = _libm_pow(TF(1.)/(TF(1.)/_libm_pow(m0, two) + TF(1.)/_libm_pow(zz, two)), TF(1.)/two)
    else:
        fac = m0
    _S(g, at)[...] -= (TF(ce1) + TF(ce2)*fac/m0)*_libm_pow(_S(g, a), TF(3./2.))/fac

def tendency_limiter(g, at, a, min_value, dt):
    """src/limiter.cxx:35-59"""
    TF = g.TF
    dt = TF(dt); mv = TF(min_value)
    dti = TF(1.)/dt
    t = _S(g, at)
    a_new = _S(g, a) + dt*t
    t[...] += np.where(a_new < mv, (-a_new + mv)*dti, TF(0.))


# --------------------------------------------------------------------------------------
# Thermo_dry (reference src/thermo_dry.cxx:66-78, 165-179)
# --------------------------------------------------------------------------------------
def thermo_dry_N2(g, N2, th, thref):
    TF = g.TF
    _S(g, N2)[...] = TF(GRAV)/_K(g, thref)*TF(0.5)*(_S(g, th, 1) - _S(g, th, -1))*_K(g, g.dzi)

def thermo_dry_buoyancy_tend_2nd(g, wt, th, threfh):
    TF = g.TF
    k0, k1 = g.kstart+1, g.kend
    _S(g, wt, 0, 0, 0, k0, k1)[...] += TF(GRAV)/_K(g, threfh, 0, k0, k1) * (
        interp2(_S(g, th, -1, 0, 0, k0, k1), _S(g, th, 0, 0, 0, k0, k1)) - _K(g, threfh, 0, k0, k1))


# --------------------------------------------------------------------------------------
# FFTW r2r semantics (reference call sites src/fft.cxx:145-155; FFTW3 manual "The Halfcomplex-format DFT")
# --------------------------------------------------------------------------------------
import scipy.fft as _sfft


def r2hc(x, axis):
    """FFTW_R2HC along `axis`: [r0, r1, ..., r_{n/2}, i_{(n+1)/2-1}, ..., i_1], unnormalised."""
    n = x.shape[axis]
    # the USESP build of the reference links fftwf: single-precision data are transformed in single precision
    # (scipy's pocketfft keeps float32; numpy.fft would silently promote to double and flatter the fp32 result)
    X = _sfft.rfft(x, axis=axis) if x.dtype == np.float32 else np.fft.rfft(x.astype(np.float64), axis=axis)
    out = np.empty(x.shape, X.real.dtype)
    sl = [slice(None)]*x.ndim
    def put(dst, src):
        d = list(sl); d[axis] = dst
        out[tuple(d)] = src
    nre = n//2 + 1
    put(slice(0, nre), X.real)
    nim = (n+1)//2 - 1
    if nim > 0:
        s = list(sl); s[axis] = slice(nim, 0, -1)
        put(slice(nre, n), X.imag[tuple(s)])
    return out.astype(x.dtype)

def hc2r(x, axis):
    """FFTW_HC2R along `axis` (unnormalised inverse of r2hc)."""
    n = x.shape[axis]
    nre = n//2 + 1
    shp = list(x.shape); shp[axis] = nre
    X = np.zeros(shp, np.complex64 if x.dtype == np.float32 else np.complex128)
    sl = [slice(None)]*x.ndim
    s = list(sl); s[axis] = slice(0, nre)
    X.real[...] = x[tuple(s)]
    nim = (n+1)//2 - 1
    if nim > 0:
        d = list(sl); d[axis] = slice(1, nim+1)
        s2 = list(sl); s2[axis] = slice(n-1, nre-1, -1)
        X.imag[tuple(d)] = x[tuple(s2)]
    out = (_sfft.irfft(X, n=n, axis=axis) if x.dtype == np.float32 else np.fft.irfft(X, n=n, axis=axis)) * x.dtype.type(n)
    return out.astype(x.dtype)


# --------------------------------------------------------------------------------------
# Pres_2 (reference src/pres_2.cxx:124-422)
# --------------------------------------------------------------------------------------
class Pres2:
    def __init__(self, g, rhoref, rhorefh):
        """set_values: src/pres_2.cxx:124-153"""
        TF = g.TF
        self.g = g
        # the reference's own precision mix (pinned bit for bit against the compiled Pres_2::set_values): `2.`, `1.` are
        # double literals, so the cosine and the products run in double from a TF-rounded pi and are narrowed on the store
        D = np.float64
        dxidxi = TF(1./D(TF(g.dx)*TF(g.dx))); dyidyi = TF(1./D(TF(g.dy)*TF(g.dy)))
        pi = D(TF(np.arccos(-1.)))
        self.bmati = np.zeros(g.itot, TF); self.bmatj = np.zeros(g.jtot, TF)
        for j in range(g.jtot//2+1):
            self.bmatj[j] = TF(2. * (np.cos(2.*pi*D(TF(j))/D(TF(g.jtot)))-1.) * D(dyidyi))
        for j in range(g.jtot//2+1, g.jtot):
            self.bmatj[j] = self.bmatj[g.jtot-j]
        for i in range(g.itot//2+1):
            self.bmati[i] = TF(2. * (np.cos(2.*pi*D(TF(i))/D(TF(g.itot)))-1.) * D(dxidxi))
        for i in range(g.itot//2+1, g.itot):
            self.bmati[i] = self.bmati[g.itot-i]
        kgc = g.kgc
        self.a = np.zeros(g.kmax, TF); self.c = np.zeros(g.kmax, TF)
        for k in range(g.kmax):
            self.a[k] = g.dz[k+kgc] * rhorefh[k+kgc]*g.dzhi[k+kgc]
            self.c[k] = g.dz[k+kgc] * rhorefh[k+kgc+1]*g.dzhi[k+kgc+1]
        self.rhoref, self.rhorefh = rhoref, rhorefh

    def input(self, u, v, w, ut, vt, wt, dt):
        """src/pres_2.cxx:155-196; returns compact p (kmax, jmax, imax)"""
        g = self.g; TF = g.TF
        dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
        dti = TF(TF(1.)/dt)       # TF(1.)/double dt, then stored in a TF
        boundary_cyclic(g, ut, EDGE_EW)
        boundary_cyclic(g, vt, EDGE_NS)
        rho = _K(g, self.rhoref)
        return ( rho * ( (_S(g, ut, 0, 0, 1) + _S(g, u, 0, 0, 1) * dti) - (_S(g, ut) + _S(g, u) * dti) ) * dxi
               + rho * ( (_S(g, vt, 0, 1, 0) + _S(g, v, 0, 1, 0) * dti) - (_S(g, vt) + _S(g, v) * dti) ) * dyi
               + ( _K(g, self.rhorefh, 1) * (_S(g, wt, 1) + _S(g, w, 1) * dti)
                 - _K(g, self.rhorefh, 0) * (_S(g, wt) + _S(g, w) * dti) ) * _K(g, g.dzi) )

    def tdma(self, p, b):
        """src/pres_2.cxx:202-263 (vectorised over the i,j plane; same per-column operation order)"""
        a, c = self.a, self.c
        kmax = self.g.kmax
        work3d = np.zeros_like(p)
        work2d = b[0].copy()
        p[0] /= work2d
        for k in range(1, kmax):
            work3d[k] = c[k-1] / work2d
            work2d = b[k] - a[k]*work3d[k]
            p[k] -= a[k]*p[k-1]
            p[k] /= work2d
        for k in range(kmax-2, -1, -1):
            p[k] -= work3d[k+1]*p[k+1]

    def solve(self, pc, p, tdma=None):
        """src/pres_2.cxx:266-362.  pc: compact rhs; p: ghosted output array."""
        g = self.g; TF = g.TF
        kgc = g.kgc
        # fft.exec_forward: x then y (src/fft.cxx:338-394)
        pc = r2hc(pc, axis=2)
        pc = r2hc(pc, axis=1)
        dz = g.dz[kgc:kgc+g.kmax][:, None, None]
        rho = self.rhoref[kgc:kgc+g.kmax][:, None, None]
        b = dz*dz * rho*(self.bmati[None, None, :] + self.bmatj[None, :, None]) - (self.a + self.c)[:, None, None]
        pc = dz*dz * pc
        b[0] += self.a[0]
        top = np.full((g.jtot, g.itot), self.c[g.kmax-1], TF)
        top[0, 0] = -self.c[g.kmax-1]
        b[g.kmax-1] += top
        pc = np.ascontiguousarray(pc); b = np.ascontiguousarray(b.astype(TF))
        (tdma or self.tdma)(pc, b)
        # fft.exec_backward: y then x with /jtot, /itot (src/fft.cxx:396-452)
        pc = hc2r(pc, axis=1) / TF(g.jtot)
        pc = hc2r(pc, axis=2) / TF(g.itot)
        _S(g, p)[...] = pc
        p[g.kstart-1, g.jstart:g.jend, g.istart:g.iend] = p[g.kstart, g.jstart:g.jend, g.istart:g.iend]
        boundary_cyclic(g, p)

    def output(self, ut, vt, wt, p):
        """src/pres_2.cxx:364-387"""
        g = self.g; TF = g.TF
        dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
        _S(g, ut)[...] -= (_S(g, p) - _S(g, p, 0, 0, -1)) * dxi
        _S(g, vt)[...] -= (_S(g, p) - _S(g, p, 0, -1, 0)) * dyi
        _S(g, wt)[...] -= (_S(g, p) - _S(g, p, -1, 0, 0)) * _K(g, g.dzhi)

    def exec(self, p, u, v, w, ut, vt, wt, dt, tdma=None):
        """src/pres_2.cxx:66-94"""
        pc = self.input(u, v, w, ut, vt, wt, dt)
        self.solve(pc, p, tdma)
        self.output(ut, vt, wt, p)

    def divergence(self, u, v, w):
        """src/pres_2.cxx:390-422"""
        g = self.g; TF = g.TF
        dxi, dyi = TF(1.)/g.dx, TF(1.)/g.dy
        div = ( _K(g, self.rhoref)*((_S(g, u, 0, 0, 1)-_S(g, u))*dxi + (_S(g, v, 0, 1, 0)-_S(g, v))*dyi)
              + (_K(g, self.rhorefh, 1)*_S(g, w, 1)-_K(g, self.rhorefh, 0)*_S(g, w))*_K(g, g.dzi) )
        return TF(np.abs(div).max())


# --------------------------------------------------------------------------------------
# Pres_4 (reference src/pres_4.cxx:178-767): 4th-order divergence / gradient (cg weights), 4-term cosine modified
# wavenumbers, a 7-band system of kmax+4 rows per horizontal mode (two boundary rows at either end) solved by
# banded LU without pivoting (`hdma`).  Member functions of the reference class (need live Grid/Fields objects), so
# restated here; PARITY UNPINNED against reference output -- pinned by the defining property instead: after
# `output` the 4th-order divergence of u + dt*ut vanishes to rounding (tests/test_pres4_oracle.py).
# --------------------------------------------------------------------------------------
class Pres4:
    def __init__(self, g):
        """set_values: src/pres_4.cxx:178-252"""
        assert g.order == 4
        TF = g.TF; D = np.float64
        self.g = g
        dxidxi = TF(1./D(g.dx*g.dx)); dyidyi = TF(1./D(g.dy*g.dy))
        pi = D(TF(np.arccos(-1.)))          # `const TF pi = std::acos(-1.)`, then used in double expressions
        def bmat(n, fac):
            b = np.zeros(n, TF)
            for q in range(n//2 + 1):
                b[q] = TF(( 2.*(1./576.)*np.cos(6.*pi*D(q)/D(n)) - 2.*(54./576.)*np.cos(4.*pi*D(q)/D(n))
                          + 2.*(783./576.)*np.cos(2.*pi*D(q)/D(n)) - (1460./576.) ) * D(fac))
            for q in range(n//2 + 1, n):
                b[q] = b[n-q]
            return b
        self.bmati = bmat(g.itot, dxidxi); self.bmatj = bmat(g.jtot, dyidyi)
        kmax, ks = g.kmax, g.kstart
        h = g.dzhi4.astype(D); c = g.dzi4.astype(D)
        m = np.zeros((7, kmax), D)
        f = 1./576.
        k, kc = 0, ks
        m[0, k] = 0.
        m[1, k] = f*(-27.*h[kc])*c[kc]
        m[2, k] = f*(-1.*h[kc+1] + 729.*h[kc] + 27.*h[kc+1])*c[kc]
        m[3, k] = f*(27.*h[kc+1] - 729.*h[kc] - 729.*h[kc+1] - 1.*h[kc+2])*c[kc]
        m[4, k] = f*(-27.*h[kc+1] + 27.*h[kc] + 729.*h[kc+1] + 27.*h[kc+2])*c[kc]
        m[5, k] = f*(1.*h[kc+1] - 27.*h[kc+1] - 27.*h[kc+2])*c[kc]
        m[6, k] = f*(1.*h[kc+2])*c[kc]
        for k in range(1, kmax-1):
            kc = ks + k
            m[0, k] = f*(1.*h[kc-1])*c[kc]
            m[1, k] = f*(-27.*h[kc-1] - 27.*h[kc])*c[kc]
            m[2, k] = f*(27.*h[kc-1] + 729.*h[kc] + 27.*h[kc+1])*c[kc]
            m[3, k] = f*(-1.*h[kc-1] - 729.*h[kc] - 729.*h[kc+1] - 1.*h[kc+2])*c[kc]
            m[4, k] = f*(27.*h[kc] + 729.*h[kc+1] + 27.*h[kc+2])*c[kc]
            m[5, k] = f*(-27.*h[kc+1] - 27.*h[kc+2])*c[kc]
            m[6, k] = f*(1.*h[kc+2])*c[kc]
        k = kmax-1; kc = ks + k
        m[0, k] = f*(1.*h[kc-1])*c[kc]
        m[1, k] = f*(-27.*h[kc-1] - 27.*h[kc] + 1.*h[kc])*c[kc]
        m[2, k] = f*(27.*h[kc-1] + 729.*h[kc] + 27.*h[kc+1] - 27.*h[kc])*c[kc]
        m[3, k] = f*(-1.*h[kc-1] - 729.*h[kc] - 729.*h[kc+1] + 27.*h[kc])*c[kc]
        m[4, k] = f*(27.*h[kc] + 729.*h[kc+1] - 1.*h[kc])*c[kc]
        m[5, k] = f*(-27.*h[kc+1])*c[kc]
        m[6, k] = 0.
        self.m = m.astype(TF)          # m1..m7 are std::vector<TF>: narrowed on assignment

    def input(self, u, v, w, ut, vt, wt, dt):
        """src/pres_4.cxx:254-317 (fills the ut/vt cyclic ghosts and the wt wall ghosts as a side effect)"""
        g = self.g; TF = g.TF
        dxi, dyi = TF(1./np.float64(g.dx)), TF(1./np.float64(g.dy))
        dti = TF(1./np.float64(TF(dt)))      # `const TF dt` narrows the sub-step first, then `1./dt` in double (src/pres_4.cxx:262,276)
        dim3 = g.jtot > 1
        boundary_cyclic(g, ut, EDGE_EW)
        if dim3:
            boundary_cyclic(g, vt, EDGE_NS)
        js, je, i0, i1, ks, ke = g.jstart, g.jend, g.istart, g.iend, g.kstart, g.kend
        wt[ks-1, js:je, i0:i1] = -wt[ks+1, js:je, i0:i1]
        wt[ke+1, js:je, i0:i1] = -wt[ke-1, js:je, i0:i1]
        cg = [TF(x) for x in CG]
        T = lambda a, at, dk=0, dj=0, di=0: _S(g, at, dk, dj, di) + _S(g, a, dk, dj, di)*dti
        p = (cg[0]*T(u, ut, 0, 0, -1) + cg[1]*T(u, ut) + cg[2]*T(u, ut, 0, 0, 1) + cg[3]*T(u, ut, 0, 0, 2)) * dxi
        if dim3:
            p = p + (cg[0]*T(v, vt, 0, -1) + cg[1]*T(v, vt) + cg[2]*T(v, vt, 0, 1) + cg[3]*T(v, vt, 0, 2)) * dyi
        p = p + (cg[0]*T(w, wt, -1) + cg[1]*T(w, wt) + cg[2]*T(w, wt, 1) + cg[3]*T(w, wt, 2)) * _K(g, g.dzi4)
        return np.ascontiguousarray(p.astype(TF))

    def hdma(self, m1, m2, m3, m4, m5, m6, m7, p):
        """src/pres_4.cxx:573-730: banded LU + forward/backward substitution; arrays are (kmax+4, ...) and modified in place"""
        TF = self.g.TF; one = TF(1.)
        kmax = self.g.kmax
        m1[0] = one; m2[0] = one; m3[0] = one/m4[0]; m4[0] = one
        m5[0] = m5[0]*m3[0]; m6[0] = m6[0]*m3[0]; m7[0] = m7[0]*m3[0]
        k = 1
        m1[k] = one; m2[k] = one
        m3[k] = m3[k]/m4[k-1]
        m4[k] = m4[k] - m3[k]*m5[k-1]
        m5[k] = m5[k] - m3[k]*m6[k-1]
        m6[k] = m6[k] - m3[k]*m7[k-1]
        k = 2
        m1[k] = one
        m2[k] = m2[k]/m4[k-2]
        m3[k] = (m3[k] - m2[k]*m5[k-2])/m4[k-1]
        m4[k] = m4[k] - m3[k]*m5[k-1] - m2[k]*m6[k-2]
        m5[k] = m5[k] - m3[k]*m6[k-1] - m2[k]*m7[k-2]
        m6[k] = m6[k] - m3[k]*m7[k-1]
        def lower(k):
            m1[k] = m1[k]/m4[k-3]
            m2[k] = (m2[k] - m1[k]*m5[k-3])/m4[k-2]
            m3[k] = (m3[k] - m2[k]*m5[k-2] - m1[k]*m6[k-3])/m4[k-1]
            m4[k] = m4[k] - m3[k]*m5[k-1] - m2[k]*m6[k-2] - m1[k]*m7[k-3]
        for k in range(3, kmax+2):
            lower(k)
            m5[k] = m5[k] - m3[k]*m6[k-1] - m2[k]*m7[k-2]
            m6[k] = m6[k] - m3[k]*m7[k-1]
        m7[kmax+1] = one
        k = kmax+2
        lower(k)
        m5[k] = m5[k] - m3[k]*m6[k-1] - m2[k]*m7[k-2]
        m6[k] = one; m7[k] = one
        k = kmax+3
        lower(k)
        m5[k] = one; m6[k] = one; m7[k] = one
        # L y = p
        p[0] = p[0]*m3[0]
        p[1] = p[1] - p[0]*m3[1]
        p[2] = p[2] - p[1]*m3[2] - p[0]*m2[2]
        for k in range(3, kmax+4):
            p[k] = p[k] - p[k-1]*m3[k] - p[k-2]*m2[k] - p[k-3]*m1[k]
        # U x = y
        k = kmax+3
        p[k] = p[k]/m4[k]
        p[k-1] = (p[k-1] - p[k]*m5[k-1])/m4[k-1]
        p[k-2] = (p[k-2] - p[k-1]*m5[k-2] - p[k]*m6[k-2])/m4[k-2]
        for k in range(kmax, -1, -1):
            p[k] = (p[k] - p[k+1]*m5[k] - p[k+2]*m6[k] - p[k+3]*m7[k])/m4[k]

    def solve(self, pc, p):
        """src/pres_4.cxx:319-529.  pc: compact rhs (kmax, jmax, imax); p: ghosted output."""
        g = self.g; TF = g.TF
        kmax = g.kmax
        pc = r2hc(pc, axis=2)
        if g.jtot > 1:
            pc = r2hc(pc, axis=1)
        shp = (kmax+4, g.jtot, g.itot)
        M = [np.zeros(shp, TF) for _ in range(7)]
        pt = np.zeros(shp, TF)
        # rows 0, 1: zero gradient at the bottom
        M[3][0] = 1.; M[6][0] = -1.
        M[3][1] = 1.; M[4][1] = -1.
        for n in range(7):
            M[n][2:kmax+2] = self.m[n][:, None, None]
        M[3][2:kmax+2] = M[3][2:kmax+2] + self.bmati[None, None, :] + self.bmatj[None, :, None]
        pt[2:kmax+2] = pc
        # top rows: dp/dz = 0, except mode (0,0) which fixes the level of p
        M[2][kmax+2] = -1.; M[3][kmax+2] = 1.
        M[0][kmax+3] = -1.; M[3][kmax+3] = 1.
        M[0][kmax+2, 0, 0] = 0.; M[1][kmax+2, 0, 0] = TF(-1/3.); M[2][kmax+2, 0, 0] = 2.; M[3][kmax+2, 0, 0] = 1.
        M[0][kmax+3, 0, 0] = -2.; M[1][kmax+3, 0, 0] = 9.; M[2][kmax+3, 0, 0] = 0.; M[3][kmax+3, 0, 0] = 1.
        self.hdma(*M, pt)
        pc = np.ascontiguousarray(pt[2:kmax+2])
        if g.jtot > 1:
            pc = hc2r(pc, axis=1) / TF(g.jtot)
        pc = hc2r(pc, axis=2) / TF(g.itot)
        _S(g, p)[...] = pc
        ks, ke = g.kstart, g.kend
        I = (slice(g.jstart, g.jend), slice(g.istart, g.iend))
        p[ks-1][I] = p[ks][I]; p[ks-2][I] = p[ks+1][I]
        p[ke][I] = p[ke-1][I]; p[ke+1][I] = p[ke-2][I]
        boundary_cyclic(g, p)

    def output(self, ut, vt, wt, p):
        """src/pres_4.cxx:531-571"""
        g = self.g; TF = g.TF
        dxi, dyi = TF(1./np.float64(g.dx)), TF(1./np.float64(g.dy))
        cg = [TF(x) for x in CG]
        P = lambda dk=0, dj=0, di=0, k0=None, k1=None: _S(g, p, dk, dj, di, k0, k1)
        _S(g, ut)[...] -= (cg[0]*P(0, 0, -2) + cg[1]*P(0, 0, -1) + cg[2]*P() + cg[3]*P(0, 0, 1)) * dxi
        if g.jtot > 1:
            _S(g, vt)[...] -= (cg[0]*P(0, -2) + cg[1]*P(0, -1) + cg[2]*P() + cg[3]*P(0, 1)) * dyi
        k0, k1 = g.kstart+1, g.kend
        _S(g, wt, 0, 0, 0, k0, k1)[...] -= (cg[0]*P(-2, 0, 0, k0, k1) + cg[1]*P(-1, 0, 0, k0, k1) + cg[2]*P(0, 0, 0, k0, k1)
                                            + cg[3]*P(1, 0, 0, k0, k1)) * _K(g, g.dzhi4, 0, k0, k1)

    def exec(self, p, u, v, w, ut, vt, wt, dt):
        """src/pres_4.cxx:76-144"""
        pc = self.input(u, v, w, ut, vt, wt, dt)
        self.solve(pc, p)
        self.output(ut, vt, wt, p)

    def divergence(self, u, v, w):
        """src/pres_4.cxx:732-767"""
        g = self.g; TF = g.TF
        dxi, dyi = TF(1./np.float64(g.dx)), TF(1./np.float64(g.dy))
        cg = [TF(x) for x in CG]
        div = ( (cg[0]*_S(g, u, 0, 0, -1) + cg[1]*_S(g, u) + cg[2]*_S(g, u, 0, 0, 1) + cg[3]*_S(g, u, 0, 0, 2)) * dxi
              + (cg[0]*_S(g, v, 0, -1) + cg[1]*_S(g, v) + cg[2]*_S(g, v, 0, 1) + cg[3]*_S(g, v, 0, 2)) * dyi
              + (cg[0]*_S(g, w, -1) + cg[1]*_S(g, w) + cg[2]*_S(g, w, 1) + cg[3]*_S(g, w, 2)) * _K(g, g.dzi4) )
        return TF(np.abs(div).max())


# --------------------------------------------------------------------------------------
# Timeloop rk3 (reference src/timeloop.cxx:250-286, 415-423)
# --------------------------------------------------------------------------------------
RK3_CA = (0., -5./9., -153./128.)
RK3_CB = (1./3., 15./16., 8./15.)

def rk3(g, a, at, substep, dt):
    TF = g.TF
    _S(g, a)[...] += TF(RK3_CB[substep])*TF(dt)*_S(g, at)
    substepn = (substep+1) % 3
    if substepn == 0:
        at[...] = TF(0.)
    else:
        _S(g, at)[...] *= TF(RK3_CA[substepn])

def rk3_subdt(dt, substep):
    """get_sub_time_step: double arithmetic (src/timeloop.cxx:338-342, 415-423)"""
    return RK3_CB[substep]*float(dt)


# --------------------------------------------------------------------------------------
# Same call surface as refbind.RefKernels, backed by the numpy restatement above
# --------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------
# Boundary_surface: Monin-Obukhov surface model (reference src/boundary_surface.cxx:55-340, 836-990;
# include/boundary_surface_kernels.h:78-330; include/monin_obukhov.h).  Pinned against the reference's own compiled kernels to
# rounding (tests/test_oracle_vs_ref.py): the transcendental functions come from numpy instead of libm and the table search is
# done at float accuracy as in the reference, so the comparison is relative 1e-12 (fp64), not bitwise.
# --------------------------------------------------------------------------------------
BC_FLUX, BC_USTAR = 2, 3
NZL_LUT = 10000          # include/boundary.h:56
ZL_MAX, ZL_MIN = 10., -1.e4          # include/constants.h:55-56
DBIG, DHUGE = 1.e9, 1.e30


def _most_psim(zeta):
    """psim_unstable / psim_stable (include/monin_obukhov.h:75-96), selected like fm() does: by the sign of L (= sign of zeta)."""
    TF = zeta.dtype.type
    with np.errstate(all="ignore"):
        phim_u = np.power(TF(1.) + TF(3.6)*np.power(np.abs(zeta), TF(2./3.)), TF(-1./2.))
        un = TF(3.)*np.log((TF(1.) + TF(1.)/phim_u)/TF(2.))
        a, b, c, d = TF(1), TF(2)/TF(3), TF(5), TF(0.35)
        st = -b*(zeta - (c/d))*np.exp(-d*zeta) - a*zeta - (b*c)/d
    return un, st


def _most_psih(zeta):
    TF = zeta.dtype.type
    with np.errstate(all="ignore"):
        phih_u = np.power(TF(1.) + TF(7.9)*np.power(np.abs(zeta), TF(2./3.)), TF(-1./2.))
        un = TF(3.)*np.log((TF(1.) + TF(1.)/phih_u)/TF(2.))
        a, b, c, d = TF(1), TF(2)/TF(3), TF(5), TF(0.35)
        st = -b*(zeta - (c/d))*np.exp(-d*zeta) - np.power(TF(1) + b*a*zeta, TF(1.5)) - (b*c)/d + TF(1)
    return un, st


def most_fm(zsl, z0m, L):
    """include/monin_obukhov.h:113-119"""
    TF = L.dtype.type
    zsl = TF(zsl); z0m = np.asarray(z0m, TF)
    u1, s1 = _most_psim(zsl/L); u0, s0 = _most_psim(z0m/L)
    lg = np.log(zsl/z0m)
    return np.where(L <= TF(0.), TF(KAPPA)/(lg - u1 + u0), TF(KAPPA)/(lg - s1 + s0)).astype(TF)


def most_fh(zsl, z0h, L):
    """include/monin_obukhov.h:121-127"""
    TF = L.dtype.type
    zsl = TF(zsl); z0h = np.asarray(z0h, TF)
    u1, s1 = _most_psih(zsl/L); u0, s0 = _most_psih(z0h/L)
    lg = np.log(zsl/z0h)
    return np.where(L <= TF(0.), TF(KAPPA)/(lg - u1 + u0), TF(KAPPA)/(lg - s1 + s0)).astype(TF)


def most_phim(zeta):
    TF = zeta.dtype.type
    with np.errstate(all="ignore"):
        un = np.power(TF(1.) + TF(3.6)*np.power(np.abs(zeta), TF(2./3.)), TF(-1./2.))
    return np.where(zeta <= TF(0.), un, TF(1) + TF(5)*zeta).astype(TF)


def most_phih(zeta):
    TF = zeta.dtype.type
    with np.errstate(all="ignore"):
        un = np.power(TF(1.) + TF(7.9)*np.power(np.abs(zeta), TF(2./3.)), TF(-1./2.))
    return np.where(zeta <= TF(0.), un, (TF(1) + TF(4)*zeta)**2).astype(TF)


def surface_prepare_lut(TF, z0m, z0h, zsl, mbcbot, thermobc, nlut=NZL_LUT):
    """bsk::prepare_lut (include/boundary_surface_kernels.h:78-138): z/L nodes (float) and the evaluation function (float)."""
    zL_tmp = np.zeros(nlut, TF)
    zLrange_min = TF(-5.)
    dzL = TF((ZL_MAX - float(zLrange_min))/(9.*nlut/10. - 1.))
    zL_tmp[0] = -TF(ZL_MAX)
    for n in range(1, 9*nlut//10):
        zL_tmp[n] = zL_tmp[n-1] + dzL
    zLend = TF(-(ZL_MIN - float(zLrange_min)))
    r = TF(1.01); r0 = TF(DHUGE)
    while abs((float(r) - float(r0))/float(r0)) > 1.e-10:
        r0 = r
        r = TF(np.power(1. - (float(zLend)/float(dzL))*(1. - float(r)), 1./(nlut/10.)))
    for n in range(9*nlut//10, nlut):
        zL_tmp[n] = zL_tmp[n-1] + dzL
        dzL = TF(dzL*r)
    zL_sl = (-zL_tmp[::-1]).astype(np.float32)
    L = (TF(zsl)/zL_sl.astype(TF)).astype(TF)
    if mbcbot == BC_DIRICHLET and thermobc == BC_FLUX:
        f_sl = (zL_sl.astype(np.float64)*np.power(most_fm(zsl, TF(z0m), L).astype(np.float64), 3)).astype(np.float32)
    elif mbcbot == BC_DIRICHLET and thermobc == BC_DIRICHLET:
        f_sl = (zL_sl.astype(np.float64)*np.power(most_fm(zsl, TF(z0m), L).astype(np.float64), 2)
                / most_fh(zsl, TF(z0h), L).astype(np.float64)).astype(np.float32)
    else:
        f_sl = np.zeros(nlut, np.float32)
    return zL_sl, f_sl


def _find_zL(TF, zL, f, Ri):
    """bsk::find_zL (include/boundary_surface_kernels.h:245-260), vectorised: the bracket search at float accuracy ends at the
    first node with f[n] > Ri (or the table ends); inside the table the result is the linear interpolation between n-1 and n."""
    nl = len(f)
    n = np.clip(np.searchsorted(f, Ri.astype(np.float32), side="right"), 0, nl - 1)
    nm = np.maximum(n - 1, 0)
    Ri = Ri.astype(np.float32)
    with np.errstate(all="ignore"):
        interp = (zL[nm] + (Ri - f[nm])/(f[n] - f[nm])*(zL[n] - zL[nm]))
    edge = (n == 0) | (n == nl - 1)
    # the reference evaluates the float expression and converts to TF
    return np.where(edge, zL[n], interp).astype(TF), n


class BoundarySurface:
    """Boundary_surface<TF> with constant z0 and the lookup solver (drycblles: mbcbot = noslip, sbcbot[th] = flux), or
    Thermo_type::Disabled (neutral).  State: ustar, obuk (jcells, icells)."""

    def __init__(self, g, z0m, z0h, mbcbot, thermobc):
        TF = g.TF
        self.g = g
        self.mbcbot, self.thermobc = mbcbot, thermobc
        self.z0m = np.full((g.jcells, g.icells), z0m, TF); self.z0h = np.full((g.jcells, g.icells), z0h, TF)
        self.ustar = np.full((g.jcells, g.icells), DSMALL, TF)      # Boundary_surface::init_surface: ustar = dsmall, obuk = dsmall
        self.obuk = np.full((g.jcells, g.icells), DSMALL, TF)
        self.zL_sl, self.f_sl = surface_prepare_lut(TF, z0m, z0h, g.z[g.kstart], mbcbot, thermobc)

    def calc_dutot(self, u, v, ubot, vbot):
        """bsk::calc_dutot (include/boundary_surface_kernels.h:140-186)"""
        g = self.g; TF = g.TF
        k = g.kstart
        U = u[k]; V = v[k]
        js, je, is_, ie = g.jstart, g.jend, g.istart, g.iend
        def S(a, dj=0, di=0):
            return a[js+dj:je+dj, is_+di:ie+di]
        h = TF(0.5)
        uf = TF(1./9)*(h*S(U, -1, -1) + S(U, -1, 0) + S(U, -1, 1) + h*S(U, -1, 2)
                       + h*S(U, 0, -1) + S(U, 0, 0) + S(U, 0, 1) + h*S(U, 0, 2)
                       + h*S(U, 1, -1) + S(U, 1, 0) + S(U, 1, 1) + h*S(U, 1, 2))
        vf = TF(1./9)*(h*S(V, -1, -1) + S(V, 0, -1) + S(V, 1, -1) + h*S(V, 2, -1)
                       + h*S(V, -1, 0) + S(V, 0, 0) + S(V, 1, 0) + h*S(V, 2, 0)
                       + h*S(V, -1, 1) + S(V, 0, 1) + S(V, 1, 1) + h*S(V, 2, 1))
        du2 = (uf - h*(S(ubot) + S(ubot, 0, 1)))**2 + (vf - h*(S(vbot) + S(vbot, 1, 0)))**2
        dutot = np.zeros((g.jcells, g.icells), TF)
        dutot[js:je, is_:ie] = np.maximum(np.power(du2, TF(0.5)), TF(1.e-1))
        boundary_cyclic_2d(g, dutot)
        return dutot

    def exec(self, c, thref=None, threfh=None, neutral=False):
        """Boundary_surface::exec (src/boundary_surface.cxx:836-990) on the case dict `c` (u, v, th and their 2-D companions
        u_bot, u_fluxbot, u_gradbot, ..., dudz_mo, dvdz_mo, dbdz_mo are updated in place)."""
        g = self.g; TF = g.TF
        k = g.kstart
        zsl = g.z[k]
        dutot = self.calc_dutot(c["u"], c["v"], c["u_bot"], c["v_bot"])
        if neutral:
            # stability_neutral (:136-180)
            if self.mbcbot == BC_USTAR:
                self.obuk[g.jstart:g.jend, g.istart:g.iend] = -TF(DBIG)
            else:
                self.obuk[...] = -TF(DBIG)
                self.ustar[...] = dutot*most_fm(zsl, self.z0m, self.obuk)
            bfluxbot = None
        else:
            th = c[c["scalars"][0]]; name = c["scalars"][0]
            # Thermo_dry::get_buoyancy_surf / get_buoyancy_fluxbot / get_db_ref (src/thermo_dry.cxx:133-162, 700-784)
            bbot = TF(GRAV)/threfh[k]*(c[name + "_bot"] - threfh[k])
            b = TF(GRAV)/thref[k]*(th[k] - thref[k])
            bfluxbot = TF(GRAV)/threfh[k]*c[name + "_fluxbot"]
            db_ref = TF(GRAV)/thref[k]*(thref[k] - threfh[k])
            if self.mbcbot == BC_USTAR and self.thermobc == BC_FLUX:
                self.obuk[...] = -self.ustar**3/(TF(KAPPA)*bfluxbot)
            elif self.mbcbot == BC_DIRICHLET and self.thermobc == BC_FLUX:
                Ri = (-TF(KAPPA)*bfluxbot*zsl/dutot**3)
                zL, _ = _find_zL(TF, self.zL_sl, self.f_sl, Ri)
                self.obuk[...] = zsl/zL
                self.ustar[...] = dutot*most_fm(zsl, self.z0m, self.obuk)
            elif self.mbcbot == BC_DIRICHLET and self.thermobc == BC_DIRICHLET:
                db = b - bbot + db_ref
                Ri = (TF(KAPPA)*db*zsl/dutot**2)
                zL, _ = _find_zL(TF, self.zL_sl, self.f_sl, Ri)
                self.obuk[...] = zsl/zL
                self.ustar[...] = dutot*most_fm(zsl, self.z0m, self.obuk)
        ustar, obuk, z0m = self.ustar, self.obuk, self.z0m
        u, v = c["u"][k], c["v"][k]
        js, je, is_, ie = g.jstart, g.jend, g.istart, g.iend
        def S(a, dj=0, di=0):
            return a[js+dj:je+dj, is_+di:ie+di]
        # surfm (:182-290), mbcbot = Dirichlet (no-slip): fluxes from the interpolated stability function
        if self.mbcbot == BC_DIRICHLET:
            sf = ustar*most_fm(zsl, z0m, obuk)
            S(c["u_fluxbot"])[...] = -(S(u) - S(c["u_bot"]))*TF(0.5)*(S(sf, 0, -1) + S(sf))
            S(c["v_fluxbot"])[...] = -(S(v) - S(c["v_bot"]))*TF(0.5)*(S(sf, -1, 0) + S(sf))
            boundary_cyclic_2d(g, c["u_fluxbot"]); boundary_cyclic_2d(g, c["v_fluxbot"])
        else:
            raise NotImplementedError("surfm with mbcbot = ustar is not restated")
        c["u_gradbot"][...] = (u - c["u_bot"])/zsl
        c["v_gradbot"][...] = (v - c["v_bot"])/zsl
        # surfs (:292-340) per scalar
        for name in c["scalars"]:
            var = c[name][k]
            bc = c.get(name + "_bcbot", self.thermobc if name == c["scalars"][0] else BC_FLUX)
            if bc == BC_DIRICHLET:
                c[name + "_fluxbot"][...] = -(var - c[name + "_bot"])*ustar*most_fh(zsl, self.z0h, obuk)
                c[name + "_gradbot"][...] = (var - c[name + "_bot"])/zsl
            elif bc == BC_FLUX:
                c[name + "_bot"][...] = c[name + "_fluxbot"]/(ustar*most_fh(zsl, self.z0h, obuk)) + var
                c[name + "_gradbot"][...] = (var - c[name + "_bot"])/zsl
        # calc_duvdz_mo / calc_dbdz_mo (include/boundary_surface_kernels.h:188-243)
        du_c = TF(0.5)*((S(u) - S(c["u_bot"])) + (S(u, 0, 1) - S(c["u_bot"], 0, 1)))
        dv_c = TF(0.5)*((S(v) - S(c["v_bot"])) + (S(v, 1, 0) - S(c["v_bot"], 1, 0)))
        fmv = most_fm(zsl, S(z0m), S(obuk))
        uflux = -du_c*S(ustar)*fmv
        vflux = -dv_c*S(ustar)*fmv
        phim = most_phim(zsl/S(obuk))
        S(c["dudz_mo"])[...] = -uflux/(TF(KAPPA)*zsl*S(ustar))*phim
        S(c["dvdz_mo"])[...] = -vflux/(TF(KAPPA)*zsl*S(ustar))*phim
        if not neutral:
            S(c["dbdz_mo"])[...] = -S(bfluxbot)/(TF(KAPPA)*zsl*S(ustar))*most_phih(zsl/S(obuk))
        return dutot


# --------------------------------------------------------------------------------------
# Buffer (damping layer, reference src/buffer.cxx:38-58, 98-124) and Force (large-scale forcings, src/force.cxx:47-272)
# --------------------------------------------------------------------------------------
def buffer_kstart(g, zstart):
    """Buffer::create (src/buffer.cxx:103-116): first full level / half level inside the damping layer."""
    ks = g.kstart + int((g.z[g.kstart:g.kend] < g.TF(zstart)).sum())
    ksh = g.kstart + int((g.zh[g.kstart:g.kend] < g.TF(zstart)).sum())
    return ks, ksh


def calc_buffer(g, at, a, abuf, z, zstart, beta, sigma, bufferkstart):
    """calc_buffer (src/buffer.cxx:38-58): at -= sigma ((z - zstart)/(zsize - zstart))^beta (a - abuf[k]) for k >= bufferkstart."""
    TF = g.TF
    zsizebuf = TF(g.zsize) - TF(zstart)
    for k in range(bufferkstart, g.kend):
        sigmaz = TF(TF(sigma)*np.power((z[k] - TF(zstart))/zsizebuf, TF(beta)))
        at[k, g.jstart:g.jend, g.istart:g.iend] -= sigmaz*(a[k, g.jstart:g.jend, g.istart:g.iend] - abuf[k])


def field_mean(g, a):
    """Field3d_operators::calc_mean (src/field3d_operators.cxx:135-155): dz-weighted mean, summed in double."""
    w = g.dz[g.kstart:g.kend].astype(np.float64)[:, None, None]
    s = (_S(g, a).astype(np.float64)*w).sum()
    return g.TF(s/(g.itot*g.jtot*float(g.zsize)))


def force_fixed_flux(g, ut, u, uflux, utrans, dt):
    """Force::exec, Fixed_flux (src/force.cxx:612-624) + enforce_fixed_flux (:65-75)"""
    TF = g.TF
    u_mean = field_mean(g, u); ut_mean = field_mean(g, ut)
    fbody = (TF(uflux) - u_mean - TF(utrans))/TF(dt) - ut_mean
    _S(g, ut)[...] += TF(fbody)
    return fbody


def force_coriolis_2nd(g, ut, vt, u, v, ug, vg, fc, ugrid, vgrid):
    """calc_coriolis_2nd (src/force.cxx:78-108)"""
    TF = g.TF
    q = TF(0.25)
    _S(g, ut)[...] += TF(fc)*(q*(_S(g, v, 0, 0, -1) + _S(g, v) + _S(g, v, 0, 1, -1) + _S(g, v, 0, 1, 0)) + TF(vgrid) - _K(g, vg))
    _S(g, vt)[...] -= TF(fc)*(q*(_S(g, u, 0, -1, 0) + _S(g, u) + _S(g, u, 0, -1, 1) + _S(g, u, 0, 0, 1)) + TF(ugrid) - _K(g, ug))


def force_ls_source(g, st, sls):
    """calc_large_scale_source (src/force.cxx:154-170)"""
    _S(g, st)[...] += _K(g, sls)


def force_wls_local(g, st, s, wls):
    """advec_wls_2nd_local (src/force.cxx:238-272): first-order upwind with the prescribed subsidence velocity"""
    for k in range(g.kstart, g.kend):
        sl = (slice(g.jstart, g.jend), slice(g.istart, g.iend))
        if wls[k] > 0.:
            st[k][sl] -= wls[k]*(s[k][sl] - s[k-1][sl])*g.dzhi[k]
        else:
            st[k][sl] -= wls[k]*(s[k+1][sl] - s[k][sl])*g.dzhi[k+1]


# --------------------------------------------------------------------------------------
# Advec_2i4 (reference src/advec_2i4.cxx:53-518) and Advec_2i62 (src/advec_2i62.cxx:59-306): 2nd-order flux divergence of
# centred interpolations -- 2i4: interp4c in all directions, falling back to interp2 on the vertical faces next to the walls
# and to no flux through the walls themselves (the reference writes those rows out one by one; `a - 0` and `-(-b)` are exact,
# so one expression with the per-face choice gives the same bits); 2i62: interp6_ws horizontally, interp2 vertically throughout.
# --------------------------------------------------------------------------------------
def _2ix_h(TF, scheme, q, axis, g, k0, k1):
    """interpolation of q to the LOWER face of every cell along a horizontal axis ('i' or 'j'); returns f(shift) giving the face
    value at cell+shift"""
    def sh(d, shift):
        return _S(g, q, 0, d + shift if axis == 'j' else 0, d + shift if axis == 'i' else 0, k0, k1)
    if scheme == "2i4":
        return lambda shift=0: _i4c(TF, sh(-2, shift), sh(-1, shift), sh(0, shift), sh(1, shift))
    return lambda shift=0: interp6_ws(sh(-3, shift), sh(-2, shift), sh(-1, shift), sh(0, shift), sh(1, shift), sh(2, shift))

def _2ix_dxdy(g, scheme):
    TF = g.TF
    if scheme == "2i4":
        return TF(1./np.float64(g.dx)), TF(1./np.float64(g.dy))        # gd.dxi = 1./gd.dx (src/grid.cxx:252-253), passed in by exec
    return TF(1.)/g.dx, TF(1.)/g.dy                                    # src/advec_2i62.cxx:130-131

def _2ix_vface(TF, scheme, q, g, k, lo, hi):
    """vertical interpolation of q to the face below level k (one level, interior columns): order by the distance to the
    walls for 2i4 (None = no flux), interp2 for 2i62.  lo / hi: first and last face that carry a flux of order >= 2 in 2i4."""
    sl = lambda dk: q[k+dk, g.jstart:g.jend, g.istart:g.iend]
    if scheme == "2i62":
        return interp2(sl(-1), sl(0))
    if k < lo or k > hi:
        return None
    if k == lo or k == hi:
        return interp2(sl(-1), sl(0))
    return _i4c(TF, sl(-2), sl(-1), sl(0), sl(1))

def _advec_2ix_cc(g, at, a, u, v, w, rhoref, rhorefh, scheme, loc):
    """tendency of a field whose vertical location is the cell centre: u (loc 'u'), v ('v') or a scalar ('s')"""
    TF = g.TF
    dxi, dyi = _2ix_dxdy(g, scheme)
    ks, ke = g.kstart, g.kend
    A = lambda dk=0, dj=0, di=0: _S(g, a, dk, dj, di)
    U = lambda dk=0, dj=0, di=0: _S(g, u, dk, dj, di)
    V = lambda dk=0, dj=0, di=0: _S(g, v, dk, dj, di)
    hx = _2ix_h(TF, scheme, a, 'i', g, ks, ke); hy = _2ix_h(TF, scheme, a, 'j', g, ks, ke)
    if loc == 'u':
        tx = -(interp2(U(), U(0,0,1))*hx(1) - interp2(U(0,0,-1), U())*hx(0))*dxi
        ty = -(interp2(V(0,1,-1), V(0,1,0))*hy(1) - interp2(V(0,0,-1), V())*hy(0))*dyi
    elif loc == 'v':
        tx = -(interp2(U(0,-1,1), U(0,0,1))*hx(1) - interp2(U(0,-1,0), U())*hx(0))*dxi
        ty = -(interp2(V(), V(0,1,0))*hy(1) - interp2(V(0,-1,0), V())*hy(0))*dyi
    else:
        tx = -(U(0,0,1)*hx(1) - U()*hx(0))*dxi
        ty = -(V(0,1,0)*hy(1) - V()*hy(0))*dyi
    for k in range(ks, ke):
        wsl = lambda kk, dj=0, di=0: w[kk, g.jstart+dj:g.jend+dj, g.istart+di:g.iend+di]
        def wf(kk):
            if loc == 'u': return interp2(wsl(kk, 0, -1), wsl(kk))
            if loc == 'v': return interp2(wsl(kk, -1, 0), wsl(kk))
            return wsl(kk)
        top = _2ix_vface(TF, scheme, a, g, k+1, ks+1, ke-1)
        bot = _2ix_vface(TF, scheme, a, g, k, ks+1, ke-1)
        if top is not None and bot is not None:
            vert = rhorefh[k+1]*wf(k+1)*top - rhorefh[k]*wf(k)*bot
        elif top is not None:
            vert = rhorefh[k+1]*wf(k+1)*top
        else:
            vert = -rhorefh[k]*wf(k)*bot
        at[k, g.jstart:g.jend, g.istart:g.iend] += tx[k-ks] + ty[k-ks] - vert/rhoref[k]*g.dzi[k]

def advec_2ix_u(g, ut, u, v, w, rhoref, rhorefh, scheme): _advec_2ix_cc(g, ut, u, u, v, w, rhoref, rhorefh, scheme, 'u')
def advec_2ix_v(g, vt, u, v, w, rhoref, rhorefh, scheme): _advec_2ix_cc(g, vt, v, u, v, w, rhoref, rhorefh, scheme, 'v')
def advec_2ix_s(g, st, s, u, v, w, rhoref, rhorefh, scheme): _advec_2ix_cc(g, st, s, u, v, w, rhoref, rhorefh, scheme, 's')

def advec_2ix_w(g, wt, u, v, w, rhoref, rhorefh, scheme):
    """advec_w (src/advec_2i4.cxx:341-416, src/advec_2i62.cxx:206-255): levels kstart+1 .. kend-1; the vertical fluxes sit at the
    cell centres, 2i4: interp2 at the lowest and the highest centre, interp4c in between"""
    TF = g.TF
    dxi, dyi = _2ix_dxdy(g, scheme)
    ks, ke = g.kstart, g.kend
    k0, k1 = ks+1, ke
    U = lambda dk=0, dj=0, di=0: _S(g, u, dk, dj, di, k0, k1)
    V = lambda dk=0, dj=0, di=0: _S(g, v, dk, dj, di, k0, k1)
    hx = _2ix_h(TF, scheme, w, 'i', g, k0, k1); hy = _2ix_h(TF, scheme, w, 'j', g, k0, k1)
    tx = -(interp2(U(-1,0,1), U(0,0,1))*hx(1) - interp2(U(-1,0,0), U())*hx(0))*dxi
    ty = -(interp2(V(-1,1,0), V(0,1,0))*hy(1) - interp2(V(-1,0,0), V())*hy(0))*dyi
    sl = lambda kk: w[kk, g.jstart:g.jend, g.istart:g.iend]
    def centre(c):
        """w interpolated to cell centre c (between faces c and c+1)"""
        if scheme == "2i62" or c == ks or c == ke-1:
            return interp2(sl(c), sl(c+1))
        return _i4c(TF, sl(c-1), sl(c), sl(c+1), sl(c+2))
    for k in range(k0, k1):
        vert = rhoref[k]*interp2(sl(k), sl(k+1))*centre(k) - rhoref[k-1]*interp2(sl(k-1), sl(k))*centre(k-1)
        wt[k, g.jstart:g.jend, g.istart:g.iend] += tx[k-k0] + ty[k-k0] - vert/rhorefh[k]*g.dzhi[k]

def advec_2ix_cfl(g, u, v, w, dt, scheme):
    """calc_cfl (src/advec_2i4.cxx:53-107, src/advec_2i62.cxx:59-102)"""
    TF = g.TF
    ks, ke = g.kstart, g.kend
    if scheme == "2i4":
        dxi, dyi = TF(1./np.float64(g.dx)), TF(1./np.float64(g.dy))     # gd.dxi, gd.dyi (src/advec_2i4.cxx:675)
        uc = _i4c(TF, _S(g, u, 0, 0, -1), _S(g, u), _S(g, u, 0, 0, 1), _S(g, u, 0, 0, 2))
        vc = _i4c(TF, _S(g, v, 0, -1), _S(g, v), _S(g, v, 0, 1), _S(g, v, 0, 2))
        wc = _i4c(TF, _S(g, w, -1), _S(g, w), _S(g, w, 1), _S(g, w, 2)).copy()
        wc[0] = interp2(_S(g, w), _S(g, w, 1))[0]; wc[-1] = interp2(_S(g, w), _S(g, w, 1))[-1]
    else:
        dxi, dyi = TF(1./np.float64(g.dx)), TF(1./np.float64(g.dy))     # `const TF dxi = 1./dx;` (:82-83)
        uc = interp6_ws(_S(g, u, 0, 0, -2), _S(g, u, 0, 0, -1), _S(g, u), _S(g, u, 0, 0, 1), _S(g, u, 0, 0, 2), _S(g, u, 0, 0, 3))
        vc = interp6_ws(_S(g, v, 0, -2), _S(g, v, 0, -1), _S(g, v), _S(g, v, 0, 1), _S(g, v, 0, 2), _S(g, v, 0, 3))
        wc = interp2(_S(g, w), _S(g, w, 1))
    cfl = (np.abs(uc)*dxi + np.abs(vc)*dyi + np.abs(wc)*_K(g, g.dzi)).max()
    return TF(cfl)*TF(dt)


# --------------------------------------------------------------------------------------
# Thermo_buoy (reference src/thermo_buoy.cxx:41-296, exec :345-391): the prognostic scalar IS the buoyancy.
# --------------------------------------------------------------------------------------
def _sin(TF, a):
    return TF(np.sin(TF(a)))      # std::sin(TF): libm sinf / sin; numpy calls the same libm scalar routine for a 0-d input

def _cos(TF, a):
    return TF(np.cos(TF(a)))

def thermo_buoy_N2(g, N2, b, bg_n2):
    """src/thermo_buoy.cxx:48-62"""
    TF = g.TF
    _S(g, N2)[...] = TF(0.5)*(_S(g, b, 1) - _S(g, b, -1))*_K(g, g.dzi) + TF(bg_n2)

def _buoy_interp_z(g, b, order, k0, k1):
    TF = g.TF
    if order == 4:
        return _i4c(TF, _S(g, b, -2, 0, 0, k0, k1), _S(g, b, -1, 0, 0, k0, k1), _S(g, b, 0, 0, 0, k0, k1), _S(g, b, 1, 0, 0, k0, k1))
    return interp2(_S(g, b, -1, 0, 0, k0, k1), _S(g, b, 0, 0, 0, k0, k1))

def thermo_buoy_tend(g, wt, b, order=2):
    """calc_buoyancy_tend_2nd / _4th (src/thermo_buoy.cxx:93-108, 166-183)"""
    k0, k1 = g.kstart+1, g.kend
    _S(g, wt, 0, 0, 0, k0, k1)[...] += _buoy_interp_z(g, b, order, k0, k1)

def thermo_buoy_tend_slope(g, ut, wt, bt, b, u, w, alpha, n2, utrans, order=2):
    """calc_buoyancy_tend_u / _w / _b, 2nd and 4th order (src/thermo_buoy.cxx:110-164, 185-246)"""
    TF = g.TF
    sa, ca = _sin(TF, alpha), _cos(TF, alpha)
    if order == 4:
        bi = _i4c(TF, _S(g, b, 0, 0, -2), _S(g, b, 0, 0, -1), _S(g, b, 0, 0, 0), _S(g, b, 0, 0, 1))
        ui = _i4c(TF, _S(g, u, 0, 0, -1), _S(g, u, 0, 0, 0), _S(g, u, 0, 0, 1), _S(g, u, 0, 0, 2))
        wi = _i4c(TF, _S(g, w, -1), _S(g, w, 0), _S(g, w, 1), _S(g, w, 2))
    else:
        bi = interp2(_S(g, b, 0, 0, -1), _S(g, b))
        ui = interp2(_S(g, u), _S(g, u, 0, 0, 1))
        wi = interp2(_S(g, w), _S(g, w, 1))
    _S(g, ut)[...] += sa*bi
    k0, k1 = g.kstart+1, g.kend
    _S(g, wt, 0, 0, 0, k0, k1)[...] += ca*_buoy_interp_z(g, b, order, k0, k1)
    _S(g, bt)[...] -= TF(n2)*(sa*(ui + TF(utrans)) + ca*wi)

def thermo_buoy_baroclinic(g, bt, v, dbdy_ls, order=2):
    """calc_baroclinic_2nd / _4th (src/thermo_buoy.cxx:248-282)"""
    TF = g.TF
    if order == 4:
        vi = _i4c(TF, _S(g, v, 0, -1), _S(g, v), _S(g, v, 0, 1), _S(g, v, 0, 2))
    else:
        vi = interp2(_S(g, v), _S(g, v, 0, 1))
    _S(g, bt)[...] -= TF(dbdy_ls)*vi

def thermo_buoy_exec(K, c, tb, order):
    """Thermo_buoy::exec (src/thermo_buoy.cxx:345-391); tb: dict(alpha, n2, utrans, swbaroclinic, dbdy_ls); scalar 0 is b"""
    b = c["scalars"][0]
    if abs(tb.get("alpha", 0.)) > 0. or abs(tb.get("n2", 0.)) > 0.:
        K.thermo_buoy_tend_slope(c["ut"], c["wt"], c[b + "t"], c[b], c["u"], c["w"], tb.get("alpha", 0.), tb.get("n2", 0.), tb.get("utrans", 0.), order)
    else:
        K.thermo_buoy_tend(c["wt"], c[b], order)
    if tb.get("swbaroclinic", False):
        K.thermo_buoy_baroclinic(c[b + "t"], c["v"], tb["dbdy_ls"], order)


# --------------------------------------------------------------------------------------
# Thermo_moist (reference src/thermo_moist.cxx, include/thermo_moist_functions.h): liquid-water potential temperature thl and
# total water qt, buoyancy through the saturation adjustment.  Constants: include/constants.h:29-39, 73-84.  Every expression
# keeps the reference's grouping (C++ evaluates a*b/c*d left to right).
# --------------------------------------------------------------------------------------
class _MC:
    """Constants::* in precision TF (`template<typename TF> constexpr TF`: derived constants are formed in TF)"""
    def __init__(self, TF):
        self.TF = TF
        self.grav, self.Rd, self.Rv, self.cp = TF(9.81), TF(287.04), TF(461.5), TF(1005)
        self.Lv, self.Lf, self.T0, self.p0 = TF(2.501e6), TF(3.337e5), TF(273.15), TF(1.e5)
        self.Ls = self.Lv + self.Lf
        self.ep = self.Rd/self.Rv
        self.c = [TF(x) for x in (+6.1121000000E+02, +4.4393067270E+01, +1.4279398448E+00, +2.6415206946E-02, +3.0291749160E-04,
                                  +2.1159987257E-06, +7.5015702516E-09, -1.5604873363E-12, -9.9726710231E-14, -4.8165754883E-17,
                                  +1.3839187032E-18)]

_MCS = {}
def _mc(TF):
    if TF not in _MCS:
        _MCS[TF] = _MC(TF)
    return _MCS[TF]

def moist_exner(TF, p):
    """exner (functions.h:151-155): pow(p/p0, Rd/cp)"""
    C = _mc(TF)
    return _libm_pow(np.atleast_1d(np.asarray(p, TF)/C.p0).astype(TF), C.Rd/C.cp).reshape(np.shape(p)).astype(TF)

def moist_esat_liq(TF, T):
    C = _mc(TF)
    x = np.minimum(np.maximum(TF(-75.), T - C.T0), TF(50.))
    r = C.c[9] + x*C.c[10]
    for n in range(8, -1, -1):
        r = C.c[n] + x*r
    return r

def moist_qsat_liq(TF, p, T):
    C = _mc(TF); es = moist_esat_liq(TF, T)
    return C.ep*es/(p - (TF(1.) - C.ep)*es)

def moist_esat_ice(TF, T):
    C = _mc(TF)
    x = np.minimum(np.maximum(TF(-100.), T - C.T0), TF(50.))
    return TF(611.15)*_libm_exp(np.atleast_1d(TF(22.452)*x/(TF(272.55) + x)).astype(TF)).reshape(np.shape(x))

def moist_qsat_ice(TF, p, T):
    C = _mc(TF); es = moist_esat_ice(TF, T)
    return C.ep*es/(p - (TF(1.) - C.ep)*es)

def moist_water_fraction(TF, T):
    C = _mc(TF)
    return np.maximum(TF(0.), np.minimum((T - TF(233.15))/(C.T0 - TF(233.15)), TF(1.)))

def moist_qsat(TF, p, T):
    a = moist_water_fraction(TF, T)
    return a*moist_qsat_liq(TF, p, T) + (TF(1.) - a)*moist_qsat_ice(TF, p, T)

def _moist_dqsatdT(TF, p, T, es, L):
    C = _mc(TF)
    den = p - es*(TF(1.) - C.ep)
    return (C.ep/den + (TF(1.) - C.ep)*C.ep*es/(den*den))*L*es/(C.Rv*(T*T))

def moist_dqsatdT_liq(TF, p, T):
    return _moist_dqsatdT(TF, p, T, moist_esat_liq(TF, T), _mc(TF).Lv)

def moist_dqsatdT_ice(TF, p, T):
    return _moist_dqsatdT(TF, p, T, moist_esat_ice(TF, T), _mc(TF).Ls)

def moist_virtual_temperature(TF, exn, thl, qt, ql, qi):
    C = _mc(TF)
    th = thl + C.Lv*ql/(C.cp*exn) + C.Ls*qi/(C.cp*exn)
    return th*(TF(1.) - (TF(1.) - C.Rv/C.Rd)*qt - C.Rv/C.Rd*(ql + qi))

def moist_buoyancy(TF, exn, thl, qt, ql, qi, thvref):
    return _mc(TF).grav*(moist_virtual_temperature(TF, exn, thl, qt, ql, qi) - thvref)/thvref

def moist_buoyancy_no_ql(TF, thl, qt, thvref):
    C = _mc(TF)
    return C.grav*(thl*(TF(1.) - (TF(1.) - C.Rv/C.Rd)*qt) - thvref)/thvref

def moist_buoyancy_flux_no_ql(TF, thl, thlflux, qt, qtflux, thvref):
    C = _mc(TF)
    return C.grav/thvref*(thlflux*(TF(1.) - (TF(1.) - C.Rv/C.Rd)*qt) - (TF(1.) - C.Rv/C.Rd)*thl*qtflux)

def moist_sat_adjust(TF, thl, qt, p, exn):
    """sat_adjust (functions.h:164-268) for arrays thl, qt at one pressure p / exner exn: returns ql, qi, t, qs.  The Newton
    loops run per point exactly as often as the scalar code does (points drop out of the active set one by one)."""
    C = _mc(TF)
    thl = np.asarray(thl, TF); qt = np.asarray(qt, TF)
    shape = thl.shape
    thl = thl.ravel(); qt = qt.ravel()
    p = TF(p); exn = TF(exn)
    tl = thl*exn
    qs = moist_qsat_liq(TF, p, tl)
    ql = np.zeros_like(tl); qi = np.zeros_like(tl); t = tl.copy()
    sat = ~((qt - qs) <= TF(0.))
    idx = np.nonzero(sat)[0]
    if idx.size:
        tls, qts = tl[idx], qt[idx]
        warm = tls >= C.T0
        tnr = tls.copy(); tnr_old = np.full_like(tls, TF(1.e9)); niter = np.zeros(tls.shape, np.int64)
        Lvcp = C.Lv/C.cp
        with np.errstate(all="ignore"):
            while True:
                act = (np.abs(tnr - tnr_old)/tnr_old > TF(1.e-5)) & (niter < 10)
                if not act.any():
                    break
                a = np.nonzero(act)[0]
                niter[a] += 1
                tn = tnr[a]; tnr_old[a] = tn
                w = warm[a]
                new = np.empty_like(tn)
                if w.any():
                    tw = tn[w]
                    q = moist_qsat_liq(TF, p, tw)
                    f = tw - tls[a][w] - Lvcp*(qts[a][w] - q)
                    fp = TF(1.) + Lvcp*moist_dqsatdT_liq(TF, p, tw)
                    new[w] = tw - f/fp
                c = ~w
                if c.any():
                    tc = tn[c]; tlc = tls[a][c]; qtc = qts[a][c]
                    q = moist_qsat(TF, p, tc)
                    aw = moist_water_fraction(TF, tc); ai = TF(1.) - aw
                    da = np.where((aw > TF(0.)) & (aw < TF(1.)), TF(0.025), TF(0.)).astype(TF)
                    dw = moist_dqsatdT_liq(TF, p, tc); di = moist_dqsatdT_ice(TF, p, tc)
                    f = (tc - tlc - aw*C.Lv/C.cp*qtc - ai*C.Ls/C.cp*qtc
                         + aw*C.Lv/C.cp*q + ai*C.Ls/C.cp*q)
                    fp = (TF(1.)
                          - da*C.Lv/C.cp*qtc + da*C.Ls/C.cp*qtc
                          + da*C.Lv/C.cp*q - da*C.Ls/C.cp*q
                          + aw*C.Lv/C.cp*dw
                          + ai*C.Ls/C.cp*di)
                    new[c] = tc - f/fp
                tnr[a] = new
        if (niter == 10).any():
            raise RuntimeError("Non-converging saturation adjustment")
        qsw = moist_qsat_liq(TF, p, tnr)
        qsc = moist_qsat(TF, p, tnr)
        aw = moist_water_fraction(TF, tnr)
        qlqi = np.maximum(TF(0.), qts - qsc)
        ql[idx] = np.where(warm, np.maximum(TF(0.), qts - qsw), aw*qlqi)
        qi[idx] = np.where(warm, TF(0.), (TF(1.) - aw)*qlqi)
        t[idx] = tnr
        qs[idx] = np.where(warm, qsw, qsc)
    return ql.reshape(shape), qi.reshape(shape), t.reshape(shape), qs.reshape(shape)

def _sa1(TF, thl, qt, p, exn):
    ql, qi, _, _ = moist_sat_adjust(TF, np.array([thl], TF), np.array([qt], TF), p, exn)
    return ql[0], qi[0]

def moist_top_and_bot(g, thl0, qt0):
    """calc_top_and_bot (src/thermo_moist.cxx:57-76): ghost values of the reference profiles"""
    TF = g.TF; ks, ke = g.kstart, g.kend
    for a in (thl0, qt0):
        s = a[ks] - g.z[ks]*(a[ks+1] - a[ks])*g.dzhi[ks+1]
        t = a[ke-1] + (g.zh[ke] - g.z[ke-1])*(a[ke-1] - a[ke-2])*g.dzhi[ke-1]
        a[ks-1] = TF(2.)*s - a[ks]
        a[ke] = TF(2.)*t - a[ke-1]

def moist_base_state(g, thlmean, qtmean, pbot):
    """calc_base_state (functions.h:271-340): hydrostatic pressure, exner, thv and density at full and half levels.
    Returns dict(pref, prefh, rhoref, rhorefh, thvref, thvrefh, exnref, exnrefh) of kcells entries (zeros where the reference
    leaves its vectors untouched)."""
    TF = g.TF; C = _mc(TF); ks, ke = g.kstart, g.kend
    z = lambda: np.zeros(g.kcells, TF)
    pref, prefh, rho, rhoh, thv, thvh, ex, exh = z(), z(), z(), z(), z(), z(), z(), z()
    e1 = lambda x: _libm_exp(np.array([x], TF))[0]
    x1 = lambda p: moist_exner(TF, np.array([p], TF))[0]
    thlsurf = TF(0.5)*(thlmean[ks-1] + thlmean[ks]); qtsurf = TF(0.5)*(qtmean[ks-1] + qtmean[ks])
    prefh[ks] = TF(pbot); exh[ks] = x1(prefh[ks])
    ql, qi = _sa1(TF, thlsurf, qtsurf, prefh[ks], exh[ks])
    thvh[ks] = moist_virtual_temperature(TF, exh[ks], thlsurf, qtsurf, ql, qi)
    rhoh[ks] = TF(pbot)/(C.Rd*exh[ks]*thvh[ks])
    pref[ks] = prefh[ks]*e1(-C.grav*g.z[ks]/(C.Rd*exh[ks]*thvh[ks]))
    for k in range(ks+1, ke+1):
        ex[k-1] = x1(pref[k-1])
        ql, qi = _sa1(TF, thlmean[k-1], qtmean[k-1], pref[k-1], ex[k-1])
        thv[k-1] = moist_virtual_temperature(TF, ex[k-1], thlmean[k-1], qtmean[k-1], ql, qi)
        rho[k-1] = pref[k-1]/(C.Rd*ex[k-1]*thv[k-1])
        prefh[k] = prefh[k-1]*e1(-C.grav*g.dz[k-1]/(C.Rd*ex[k-1]*thv[k-1]))
        exh[k] = x1(prefh[k])
        thli = TF(0.5)*(thlmean[k-1] + thlmean[k]); qti = TF(0.5)*(qtmean[k-1] + qtmean[k])
        ql, qi = _sa1(TF, thli, qti, prefh[k], exh[k])
        thvh[k] = moist_virtual_temperature(TF, exh[k], thli, qti, ql, qi)
        rhoh[k] = prefh[k]/(C.Rd*exh[k]*thvh[k])
        pref[k] = pref[k-1]*e1(-C.grav*g.dzh[k]/(C.Rd*exh[k]*thvh[k]))
    pref[ks-1] = TF(2.)*prefh[ks] - pref[ks]
    return dict(pref=pref, prefh=prefh, rhoref=rho, rhorefh=rhoh, thvref=thv, thvrefh=thvh, exnref=ex, exnrefh=exh)

def mean_profile(g, fld):
    """Field3d_operators::calc_mean_profile (src/field3d_operators.cxx:45-66): double accumulation over the interior of every
    level (ghost levels included), in i-fastest order, divided by itot*jtot"""
    TF = g.TF
    n = np.float64(g.itot*g.jtot)
    inner = fld[:, g.jstart:g.jend, g.istart:g.iend].astype(np.float64).reshape(g.kcells, -1)
    out = np.empty(g.kcells, np.float64)
    for k in range(g.kcells):
        out[k] = np.add.accumulate(inner[k])[-1] if inner.shape[1] else 0.      # sequential sum, not numpy's pairwise one
    return (out/n).astype(TF)

def thermo_moist_buoyancy_tend_2nd(g, wt, thl, qt, ph, thvrefh):
    """calc_buoyancy_tend_2nd (src/thermo_moist.cxx:77-120)"""
    TF = g.TF
    for k in range(g.kstart+1, g.kend):
        exnh = moist_exner(TF, np.array([ph[k]], TF))[0]
        sl = (k, slice(g.jstart, g.jend), slice(g.istart, g.iend)); sm = (k-1,) + sl[1:]
        thlh = interp2(thl[sm], thl[sl]); qth = interp2(qt[sm], qt[sl])
        ql, qi, _, _ = moist_sat_adjust(TF, thlh, qth, ph[k], exnh)
        wt[sl] += moist_buoyancy(TF, exnh, thlh, qth, ql, qi, thvrefh[k])

def thermo_moist_buoyancy(g, b, thl, qt, p, thvref):
    """calc_buoyancy (src/thermo_moist.cxx:122-168): get_thermo_field("b"); no condensate outside kstart..kend-1"""
    TF = g.TF
    for k in range(g.kcells):
        ex = moist_exner(TF, np.array([p[k]], TF))[0]
        sl = (k, slice(g.jstart, g.jend), slice(g.istart, g.iend))
        if g.kstart <= k < g.kend:
            ql, qi, _, _ = moist_sat_adjust(TF, thl[sl], qt[sl], p[k], ex)
        else:
            ql = np.zeros_like(thl[sl]); qi = ql
        b[sl] = moist_buoyancy(TF, ex, thl[sl], qt[sl], ql, qi, thvref[k])

def thermo_moist_liquid_water(g, ql, thl, qt, p):
    """calc_liquid_water (src/thermo_moist.cxx:230-250): get_thermo_field("ql")"""
    TF = g.TF
    for k in range(g.kstart, g.kend):
        ex = moist_exner(TF, np.array([p[k]], TF))[0]
        sl = (k, slice(g.jstart, g.jend), slice(g.istart, g.iend))
        ql[sl] = moist_sat_adjust(TF, thl[sl], qt[sl], p[k], ex)[0]

def thermo_moist_N2(g, N2, thl, thvref):
    """calc_N2 (src/thermo_moist.cxx:459-475)"""
    TF = g.TF
    _S(g, N2)[...] = TF(GRAV)/_K(g, thvref)*TF(0.5)*(_S(g, thl, 1) - _S(g, thl, -1))*_K(g, g.dzi)

def thermo_moist_buoyancy_bot(g, b, bbot, thl, thlbot, qt, qtbot, thvref, thvrefh):
    """calc_buoyancy_bot (src/thermo_moist.cxx:637-655): whole 2-D planes, ghost cells included"""
    TF = g.TF; ks = g.kstart
    bbot[...] = moist_buoyancy_no_ql(TF, thlbot, qtbot, thvrefh[ks])
    b[ks] = moist_buoyancy_no_ql(TF, thl[ks], qt[ks], thvref[ks])

def thermo_moist_buoyancy_fluxbot(g, bfluxbot, thl, thlfluxbot, qt, qtfluxbot, thvrefh):
    """calc_buoyancy_fluxbot (src/thermo_moist.cxx:675-693)"""
    TF = g.TF; ks = g.kstart
    bfluxbot[...] = moist_buoyancy_flux_no_ql(TF, thl[ks], thlfluxbot, qt[ks], qtfluxbot, thvrefh[ks])


# --------------------------------------------------------------------------------------
# Restart IO of one 3-D field: Field3d_io<TF>::save_field3d / load_field3d (serial build, src/field3d_io.cxx:669-751; called
# for every prognostic field by Fields::save / load with offset 0, src/fields.cxx:1243-1320).  File = the interior
# [kstart,kend) x jtot x itot as raw TF in C order, no header.  The MPI build writes the same single file through MPI-IO
# subarrays (src/field3d_io.cxx:57-160): a y slab owns rows [mpicoordy*jmax, (mpicoordy+1)*jmax) of every level.
# --------------------------------------------------------------------------------------
def field3d_save(g, data, filename, offset=0., kstart=None, kend=None):
    TF = g.TF
    k0 = g.kstart if kstart is None else kstart
    k1 = g.kend if kend is None else kend
    import os
    if os.path.exists(filename):
        return 1                                    # fopen(filename, "wbx"): exclusive create
    (data[k0:k1, g.jstart:g.jend, g.istart:g.iend] + TF(offset)).astype(TF).tofile(filename)
    return 0

def field3d_load(g, data, filename, offset=0., kstart=None, kend=None):
    TF = g.TF
    k0 = g.kstart if kstart is None else kstart
    k1 = g.kend if kend is None else kend
    import os
    n = (k1 - k0)*g.jmax*g.imax
    if not os.path.exists(filename) or os.path.getsize(filename) < n*np.dtype(TF).itemsize:
        return 1
    a = np.fromfile(filename, dtype=TF, count=n).reshape(k1 - k0, g.jmax, g.imax)
    data[k0:k1, g.jstart:g.jend, g.istart:g.iend] = a - TF(offset)
    return 0


class NumpyKernels:
    def __init__(self, g):
        self.g = g

    def boundary_cyclic(self, a, edge=EDGE_BOTH): boundary_cyclic(self.g, a, edge)
    def ghost_cells_bot_2nd(self, a, bc, abot, agradbot): ghost_cells_bot_2nd(self.g, a, bc, abot, agradbot)
    def ghost_cells_top_2nd(self, a, bc, atop, agradtop): ghost_cells_top_2nd(self.g, a, bc, atop, agradtop)
    def ghost_cells_bot_4th(self, a, bc, abot, agradbot): ghost_cells_bot_4th(self.g, a, bc, abot, agradbot)
    def ghost_cells_top_4th(self, a, bc, atop, agradtop): ghost_cells_top_4th(self.g, a, bc, atop, agradtop)
    def ghost_cells_w_4th(self, w, conservation): ghost_cells_w_4th(self.g, w, conservation)
    def advec_2i5_u(self, ut, u, v, w, rhoref, rhorefh): advec_2i5_u(self.g, ut, u, v, w, rhoref, rhorefh)
    def advec_2i5_v(self, vt, u, v, w, rhoref, rhorefh): advec_2i5_v(self.g, vt, u, v, w, rhoref, rhorefh)
    def advec_2i5_w(self, wt, u, v, w, rhoref, rhorefh): advec_2i5_w(self.g, wt, u, v, w, rhoref, rhorefh)
    def advec_2i5_s(self, st, s, u, v, w, rhoref, rhorefh): advec_2i5_s(self.g, st, s, u, v, w, rhoref, rhorefh)
    def advec_2i5_cfl(self, u, v, w, dt): return float(advec_2i5_cfl(self.g, u, v, w, dt))
    def advec_s_lim(self, st, s, u, v, w, rhoref, rhorefh): advec_s_lim(self.g, st, s, u, v, w, rhoref, rhorefh)
    def advec_2_u(self, ut, u, v, w, rhoref, rhorefh): advec_2_u(self.g, ut, u, v, w, rhoref, rhorefh)
    def advec_2_v(self, vt, u, v, w, rhoref, rhorefh): advec_2_v(self.g, vt, u, v, w, rhoref, rhorefh)
    def advec_2_w(self, wt, u, v, w, rhoref, rhorefh): advec_2_w(self.g, wt, u, v, w, rhoref, rhorefh)
    def advec_2_s(self, st, s, u, v, w, rhoref, rhorefh): advec_2_s(self.g, st, s, u, v, w, rhoref, rhorefh)
    def advec_2_cfl(self, u, v, w, dt): return float(advec_2_cfl(self.g, u, v, w, dt))
    def diff_2_c(self, at, a, visc): diff_2_c(self.g, at, a, visc)
    def advec_4_u(self, ut, u, v, w): advec_4_u(self.g, ut, u, v, w)
    def advec_4_v(self, vt, u, v, w): advec_4_v(self.g, vt, u, v, w)
    def advec_4_w(self, wt, u, v, w): advec_4_w(self.g, wt, u, v, w)
    def advec_4_s(self, st, s, u, v, w): advec_4_s(self.g, st, s, u, v, w)
    def advec_4_cfl(self, u, v, w, dt): return float(advec_4_cfl(self.g, u, v, w, dt))
    def advec_4m_cfl(self, u, v, w, dt): return float(advec_4m_cfl(self.g, u, v, w, dt))
    def advec_4m_u(self, ut, u, v, w): advec_4m_u(self.g, ut, u, v, w)
    def advec_4m_v(self, vt, u, v, w): advec_4m_v(self.g, vt, u, v, w)
    def advec_4m_w(self, wt, u, v, w): advec_4m_w(self.g, wt, u, v, w)
    def advec_4m_s(self, st, s, u, v, w): advec_4m_s(self.g, st, s, u, v, w)
    def diff_4_c(self, at, a, visc): diff_4_c(self.g, at, a, visc)
    def diff_4_w(self, wt, w, visc): diff_4_w(self.g, wt, w, visc)
    def diff_2_w(self, wt, w, visc): diff_2_w(self.g, wt, w, visc)
    def diff_strain2(self, strain2, u, v, w, ugradbot, vgradbot, surface): diff_strain2(self.g, strain2, u, v, w, ugradbot, vgradbot, surface)
    def diff_evisc(self, evisc, u, v, w, N2, bgradbot, z0m, cs, tPr, surface, mason=True): diff_evisc(self.g, evisc, N2, bgradbot, z0m, cs, tPr, surface, mason)
    def diff_evisc_neutral(self, evisc, u, v, w, z0m, cs, visc, surface, mason=True): diff_evisc_neutral(self.g, evisc, u, v, w, z0m, cs, visc, surface, mason)
    def diff_u(self, ut, u, v, w, evisc, fluxbot, fluxtop, rhoref, rhorefh, visc, surface): diff_u(self.g, ut, u, v, w, evisc, fluxbot, fluxtop, rhoref, rhorefh, visc, surface)
    def diff_v(self, vt, u, v, w, evisc, fluxbot, fluxtop, rhoref, rhorefh, visc, surface): diff_v(self.g, vt, u, v, w, evisc, fluxbot, fluxtop, rhoref, rhorefh, visc, surface)
    def diff_w(self, wt, u, v, w, evisc, rhoref, rhorefh, visc): diff_w(self.g, wt, u, v, w, evisc, rhoref, rhorefh, visc)
    def diff_c(self, at, a, evisc, fluxbot, fluxtop, rhoref, rhorefh, tPr, visc, surface): diff_c(self.g, at, a, evisc, fluxbot, fluxtop, rhoref, rhorefh, tPr, visc, surface)
    def diff_dnmul(self, evisc, tPr): return float(diff_dnmul(self.g, evisc, tPr))
    def thermo_buoy_N2(self, N2, b, bg_n2): thermo_buoy_N2(self.g, N2, b, bg_n2)
    def advec_2i4_u(self, at, u, v, w, rhoref, rhorefh): advec_2ix_u(self.g, at, u, v, w, rhoref, rhorefh, "2i4")
    def advec_2i4_v(self, at, u, v, w, rhoref, rhorefh): advec_2ix_v(self.g, at, u, v, w, rhoref, rhorefh, "2i4")
    def advec_2i4_w(self, at, u, v, w, rhoref, rhorefh): advec_2ix_w(self.g, at, u, v, w, rhoref, rhorefh, "2i4")
    def advec_2i4_s(self, st, s, u, v, w, rhoref, rhorefh): advec_2ix_s(self.g, st, s, u, v, w, rhoref, rhorefh, "2i4")
    def advec_2i4_cfl(self, u, v, w, dt): return float(advec_2ix_cfl(self.g, u, v, w, dt, "2i4"))
    def advec_2i62_u(self, at, u, v, w, rhoref, rhorefh): advec_2ix_u(self.g, at, u, v, w, rhoref, rhorefh, "2i62")
    def advec_2i62_v(self, at, u, v, w, rhoref, rhorefh): advec_2ix_v(self.g, at, u, v, w, rhoref, rhorefh, "2i62")
    def advec_2i62_w(self, at, u, v, w, rhoref, rhorefh): advec_2ix_w(self.g, at, u, v, w, rhoref, rhorefh, "2i62")
    def advec_2i62_s(self, st, s, u, v, w, rhoref, rhorefh): advec_2ix_s(self.g, st, s, u, v, w, rhoref, rhorefh, "2i62")
    def advec_2i62_cfl(self, u, v, w, dt): return float(advec_2ix_cfl(self.g, u, v, w, dt, "2i62"))
    def moist_base_state(self, thlmean, qtmean, pbot): return moist_base_state(self.g, thlmean, qtmean, pbot)
    def moist_top_and_bot(self, thl0, qt0): moist_top_and_bot(self.g, thl0, qt0)
    def mean_profile(self, fld): return mean_profile(self.g, fld)
    def thermo_moist_buoyancy_tend_2nd(self, wt, thl, qt, ph, thvrefh): thermo_moist_buoyancy_tend_2nd(self.g, wt, thl, qt, ph, thvrefh)
    def thermo_moist_buoyancy(self, b, thl, qt, p, thvref): thermo_moist_buoyancy(self.g, b, thl, qt, p, thvref)
    def thermo_moist_liquid_water(self, ql, thl, qt, p): thermo_moist_liquid_water(self.g, ql, thl, qt, p)
    def thermo_moist_N2(self, N2, thl, thvref): thermo_moist_N2(self.g, N2, thl, thvref)
    def thermo_moist_buoyancy_bot(self, b, bbot, thl, thlbot, qt, qtbot, thvref, thvrefh): thermo_moist_buoyancy_bot(self.g, b, bbot, thl, thlbot, qt, qtbot, thvref, thvrefh)
    def thermo_moist_buoyancy_fluxbot(self, bfluxbot, thl, thlfluxbot, qt, qtfluxbot, thvrefh): thermo_moist_buoyancy_fluxbot(self.g, bfluxbot, thl, thlfluxbot, qt, qtfluxbot, thvrefh)
    def thermo_buoy_tend(self, wt, b, order=2): thermo_buoy_tend(self.g, wt, b, order)
    def thermo_buoy_tend_slope(self, ut, wt, bt, b, u, w, alpha, n2, utrans, order=2): thermo_buoy_tend_slope(self.g, ut, wt, bt, b, u, w, alpha, n2, utrans, order)
    def thermo_buoy_baroclinic(self, bt, v, dbdy_ls, order=2): thermo_buoy_baroclinic(self.g, bt, v, dbdy_ls, order)
    def tke2_enforce_min(self, sgstke): tke2_enforce_min(self.g, sgstke)
    def tke2_evisc_neutral(self, evisc, sgstke, u, v, w, z0m, cn, cm, mason=True): tke2_evisc_neutral(self.g, evisc, sgstke, z0m, cn, cm, mason)
    def tke2_evisc(self, evisc, sgstke, u, v, w, N2, bgradbot, z0m, cn, cm, mason=True): tke2_evisc(self.g, evisc, sgstke, N2, bgradbot, z0m, cn, cm, mason)
    def tke2_evisc_heat(self, evisch, evisc, sgstke, N2, bgradbot, z0m, cn, ch1, ch2, mason=True): tke2_evisc_heat(self.g, evisch, evisc, sgstke, N2, bgradbot, z0m, cn, ch1, ch2, mason)
    def tke2_shear_tend(self, at, a, evisc, strain2): tke2_shear_tend(self.g, at, evisc, strain2)
    def tke2_buoy_tend(self, at, a, evisch, N2, bgradbot): tke2_buoy_tend(self.g, at, evisch, N2, bgradbot)
    def tke2_diss_tend(self, at, a, N2, bgradbot, z0m, cn, ce1, ce2, mason=True): tke2_diss_tend(self.g, at, a, N2, bgradbot, z0m, cn, ce1, ce2, mason)
    def tke2_diss_tend_neutral(self, at, a, z0m, ce1, ce2, mason=True): tke2_diss_tend_neutral(self.g, at, a, z0m, ce1, ce2, mason)
    def tendency_limiter(self, at, a, min_value, dt): tendency_limiter(self.g, at, a, min_value, dt)
    def thermo_dry_N2(self, N2, th, thref): thermo_dry_N2(self.g, N2, th, thref)
    def thermo_dry_buoyancy_tend_2nd(self, wt, th, threfh): thermo_dry_buoyancy_tend_2nd(self.g, wt, th, threfh)
    def rk3(self, a, at, substep, dt): rk3(self.g, a, at, substep, dt)
    tdma = None   # Pres2.tdma is used
