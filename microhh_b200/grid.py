"""
Host-side mirror of the reference's `Grid_data<TF>` for the hot path
(reference include/grid.h:49-134; sizes src/grid.cxx:141-170; 2nd-order metric arrays
src/grid.cxx:274-304).  The C ABI takes these scalars and 1-D metric arrays from its caller --
inside MicroHH that caller is `Grid<TF>`; tests and bench.py use this class.
"""
import numpy as np


class GridData:
    def __init__(self, itot, jtot, ktot, xsize, ysize, zsize, igc=3, jgc=3, kgc=1,
                 dtype=np.float64, z=None, npx=1, npy=1, mpicoordx=0, mpicoordy=0):
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
        dzi[1:kc-1] = TF(1.)/dz[1:kc-1]
        dz[ks-1] = dz[ks]; dzi[ks-1] = dzi[ks]
        dz[ke] = dz[ke-1]; dzi[ke] = dzi[ke-1]
        self.z, self.zh, self.dz, self.dzh, self.dzi, self.dzhi = zf, zh, dz, dzh, dzi, dzhi

    @property
    def shape(self):
        return (self.kcells, self.jcells, self.icells)

    @property
    def shape2d(self):
        return (self.jcells, self.icells)

    @property
    def npoints(self):
        return self.itot*self.jtot*self.ktot
