"""
Host-side mirror of the reference's `Grid_data<TF>` for the hot path
(reference include/grid.h:49-134; sizes src/grid.cxx:141-170; 2nd-order metric arrays
src/grid.cxx:274-304).  The C ABI takes these scalars and 1-D metric arrays from its caller --
inside MicroHH that caller is `Grid<TF>`; tests and bench.py use this class.
"""
import numpy as np


class GridData:
    def __init__(self, itot, jtot, ktot, xsize, ysize, zsize, igc=3, jgc=3, kgc=1,
                 dtype=np.float64, z=None, npx=1, npy=1, mpicoordx=0, mpicoordy=0, order=2):
        TF = np.dtype(dtype).type
        self.TF = TF
        self.dtype = np.dtype(dtype)
        self.itot, self.jtot, self.ktot = itot, jtot, ktot
        self.npx, self.npy = npx, npy
        self.mpicoordx, self.mpicoordy = mpicoordx, mpicoordy
        # local block (src/grid.cxx:142-150)
        self.imax, self.jmax, self.kmax = itot // npx, jtot // npy, ktot
        self.igc, self.jgc, self.kgc = igc, jgc, kgc
        self.icells = self.imax + 2*igc
        self.jcells = self.jmax + 2*jgc
        self.kcells = self.kmax + 2*kgc
        self.ijcells = self.icells*self.jcells
        self.ncells = self.ijcells*self.kcells
        self.istart, self.iend = igc, igc + self.imax
        self.jstart, self.jend = jgc, jgc + self.jmax
        self.kstart, self.kend = kgc, kgc + self.kmax
        self.xsize, self.ysize, self.zsize = TF(xsize), TF(ysize), TF(zsize)
        self.dx = TF(self.xsize / itot)
        self.dy = TF(self.ysize / jtot)

        ks, ke, kc = self.kstart, self.kend, self.kcells
        self.order = order
        if order == 4:
            if min(igc, jgc, kgc) < 3:
                raise ValueError("swspatialorder=4 needs three ghost cells in every direction (src/grid.cxx:87-92)")
            self._metrics_4th(z)
            return
        if z is None:
            dz0 = zsize / ktot
            z = np.linspace(0.5*dz0, zsize - 0.5*dz0, ktot)
        zf = np.zeros(kc, TF)
        zf[ks:ke] = np.asarray(z, TF)
        zf[ks-1] = -zf[ks]
        zf[ke] = TF(2.)*self.zsize - zf[ke-1]
        zh = np.zeros(kc, TF)
        zh[ks+1:ke] = TF(0.5)*(zf[ks:ke-1] + zf[ks+1:ke])
        zh[ks] = TF(0.)
        zh[ke] = self.zsize
        dzh = np.zeros(kc, TF); dzhi = np.zeros(kc, TF)
        dzh[1:] = zf[1:] - zf[:-1]
        dzhi[1:] = TF(1.)/dzh[1:]
        dzh[ks-1] = dzh[ks+1]; dzhi[ks-1] = dzhi[ks+1]
        dz = np.zeros(kc, TF); dzi = np.zeros(kc, TF)
        dz[1:kc-1] = zh[2:kc] - zh[1:kc-1]
        with np.errstate(divide="ignore"):      # kgc = 2: the outer ghost level has dz = 0 in the reference too (1/0 = inf, never read)
            dzi[1:kc-1] = TF(1.)/dz[1:kc-1]
        dz[ks-1] = dz[ks]; dzi[ks-1] = dzi[ks]
        dz[ke] = dz[ke-1]; dzi[ke] = dzi[ke-1]
        self.z, self.zh, self.dz, self.dzh, self.dzi, self.dzhi = zf, zh, dz, dzh, dzi, dzhi

    def _metrics_4th(self, z):
        """4th-order vertical metrics (src/grid.cxx:306-375): ghost heights by extrapolation through the walls, face
        heights by 4th-order interpolation, dzi4 / dzhi4 = reciprocal 4th-order gradients of the face / centre
        heights with one-sided stencils next to the walls.  Lines with double literals in the reference are
        evaluated in double and narrowed; the weighted sums of TF values stay in TF."""
        TF = self.TF
        f8 = np.float64
        ks, ke, kc = self.kstart, self.kend, self.kcells
        if z is None:
            dz0 = float(self.zsize)/self.ktot
            z = np.linspace(0.5*dz0, float(self.zsize) - 0.5*dz0, self.ktot)
        huge = 1e30
        c_i = np.array([-1, 9, 9, -1], f8)/16; b_i = np.array([5, 15, -5, 1], f8)/16; t_i = np.array([1, -5, 15, 5], f8)/16
        c_g = np.array([1, -27, 27, -1], f8)/24; b_g = np.array([-23, 21, 3, -1], f8)/24; t_g = np.array([1, -3, -21, 23], f8)/24

        def wsum(wts, arr, k0):        # TF arithmetic, left to right
            acc = TF(wts[0])*arr[k0]
            for n in range(1, 4):
                acc = acc + TF(wts[n])*arr[k0 + n]
            return acc

        zc = np.zeros(kc, TF); zf = np.zeros(kc, TF)
        zc[ks:ke] = np.asarray(z, TF)
        zc[ks-1] = TF(-2.*f8(zc[ks]) + (1./3.)*f8(zc[ks+1]))
        zc[ks-2] = TF(-9.*f8(zc[ks]) + 2.*f8(zc[ks+1]))
        zc[ke] = TF((8./3.)*f8(self.zsize) - 2.*f8(zc[ke-1]) + (1./3.)*f8(zc[ke-2]))
        zc[ke+1] = TF(8.*f8(self.zsize) - 9.*f8(zc[ke-1]) + 2.*f8(zc[ke-2]))
        zc[ks-3] = TF(huge); zc[ke+2] = TF(huge)
        for k in range(ks+1, ke):
            zf[k] = wsum(c_i, zc, k-2)
        zf[ks] = TF(0.); zf[ke] = self.zsize
        zf[ks-1] = wsum(b_i, zc, ks-2)
        zf[ke+1] = wsum(t_i, zc, ke-2)
        dzh = np.zeros(kc, TF); dzhi = np.zeros(kc, TF); dz = np.zeros(kc, TF); dzi = np.zeros(kc, TF)
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            dzh[1:] = zc[1:] - zc[:-1]
            dzhi[1:] = (1./dzh[1:].astype(f8)).astype(TF)
            dzh[ks-3] = dzh[ks+3]; dzhi[ks-3] = dzhi[ks+3]
            dz[1:kc-1] = zf[2:kc] - zf[1:kc-1]
            dzi[1:kc-1] = (1./dz[1:kc-1].astype(f8)).astype(TF)
        dz[ks-3] = dz[ks+2]; dzi[ks-3] = dzi[ks+2]
        dz[ke+2] = dz[ke-3]; dzi[ke+2] = dzi[ke-3]
        dzi4 = np.zeros(kc, TF); dzhi4 = np.zeros(kc, TF)
        rcp = lambda v: TF(1./f8(v))
        for k in range(ks, ke):
            dzi4[k] = rcp(wsum(c_g, zf, k-1))
            dzhi4[k] = rcp(wsum(c_g, zc, k-2))
        dzhi4[ke] = rcp(wsum(c_g, zc, ke-2))
        dzi4[ks-1] = rcp(wsum(b_g, zf, ks-1)); dzhi4[ks-1] = rcp(wsum(b_g, zc, ks-2))
        dzi4[ke] = rcp(wsum(t_g, zf, ke-2)); dzhi4[ke+1] = rcp(wsum(t_g, zc, ke-2))
        self.dzhi4bot = rcp(wsum(b_g, zc, ks-1)); self.dzhi4top = rcp(wsum(t_g, zc, ke-3))
        for k in (ks-3, ks-2, ke+1, ke+2):
            dzi4[k] = TF(huge)
        self.z, self.zh, self.dz, self.dzh, self.dzi, self.dzhi = zc, zf, dz, dzh, dzi, dzhi
        self.dzi4, self.dzhi4 = dzi4, dzhi4

    @property
    def shape(self):
        return (self.kcells, self.jcells, self.icells)

    @property
    def shape2d(self):
        return (self.jcells, self.icells)

    @property
    def npoints(self):
        return self.itot*self.jtot*self.ktot
