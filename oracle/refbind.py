"""
TEST INFRASTRUCTURE ONLY -- ctypes view of `oracle/_ref/libmhh_ref.so`, the reference's own
CPU kernels compiled from /root/reference by `oracle/Makefile` (see oracle/ref/*.cpp).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg use it.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

def lib_path(fast=False):
    return os.path.join(_HERE, "_ref", "libmhh_ref_fast.so" if fast else "libmhh_ref.so")

def available(fast=False):
    return os.path.exists(lib_path(fast))


class RefKernels:
    """Calls the reference kernels on numpy arrays laid out (kcells, jcells, icells)."""

    def __init__(self, g, fast=False):
        self.lib = C.CDLL(lib_path(fast))
        self.g = g
        self.sfx = "f64" if g.TF == np.float64 else "f32"
        self.ct = C.c_double if g.TF == np.float64 else C.c_float
        self.lib.ref_set_geom(g.itot, g.jtot, g.ktot, g.igc, g.jgc, g.kgc)

    def set_geom(self):
        g = self.g
        self.lib.ref_set_geom(g.itot, g.jtot, g.ktot, g.igc, g.jgc, g.kgc)

    def _p(self, a):
        if a is None:
            return None
        assert a.flags["C_CONTIGUOUS"] and a.dtype == self.g.TF, (a.dtype, a.flags)
        return a.ctypes.data_as(C.c_void_p)

    def _call(self, name, *args, restype=None):
        f = getattr(self.lib, f"{name}_{self.sfx}")
        f.restype = restype
        conv = []
        for a in args:
            if isinstance(a, np.ndarray) or a is None:
                conv.append(self._p(a))
            elif isinstance(a, (int, np.integer)):
                conv.append(C.c_int(int(a)))
            else:
                conv.append(self.ct(float(a)))
        return f(*conv)

    # --- boundary
    def boundary_cyclic(self, a, edge=2):
        self._call("ref_boundary_cyclic", a, edge)

    def ghost_cells_bot_2nd(self, a, bc, abot, agradbot):
        self._call("ref_ghost_cells_bot_2nd", a, self.g.dzh, bc, abot, agradbot)

    def ghost_cells_top_2nd(self, a, bc, atop, agradtop):
        self._call("ref_ghost_cells_top_2nd", a, self.g.dzh, bc, atop, agradtop)

    def ghost_cells_bot_4th(self, a, bc, abot, agradbot):
        self._call("ref_ghost_cells_bot_4th", a, self.g.z, bc, abot, agradbot)

    def ghost_cells_top_4th(self, a, bc, atop, agradtop):
        self._call("ref_ghost_cells_top_4th", a, self.g.z, bc, atop, agradtop)

    def ghost_cells_w_4th(self, w, conservation):
        self._call("ref_ghost_cells_w_4th", w, int(bool(conservation)))

    # --- advec_2i5
    def advec_2i5_u(self, ut, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_2i5_u", ut, u, v, w, g.dzi, g.dx, g.dy, rhoref, rhorefh)

    def advec_2i5_v(self, vt, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_2i5_v", vt, u, v, w, g.dzi, g.dx, g.dy, rhoref, rhorefh)

    def advec_2i5_w(self, wt, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_2i5_w", wt, u, v, w, g.dzhi, g.dx, g.dy, rhoref, rhorefh)

    def advec_2i5_s(self, st, s, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_2i5_s", st, s, u, v, w, g.dzi, g.dx, g.dy, rhoref, rhorefh)

    def advec_s_lim(self, st, s, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_s_lim", st, s, u, v, w, g.dzi, g.dx, g.dy, rhoref, rhorefh)

    def advec_2i5_cfl(self, u, v, w, dt):
        g = self.g
        return self._call("ref_advec_2i5_cfl", u, v, w, g.dzi, g.dx, g.dy, float(dt), restype=C.c_double)

    # --- advec_2
    def advec_2_u(self, ut, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_2_u", ut, u, v, w, g.dzi, g.dx, g.dy, rhoref, rhorefh)

    def advec_2_v(self, vt, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_2_v", vt, u, v, w, g.dzi, g.dx, g.dy, rhoref, rhorefh)

    def advec_2_w(self, wt, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_2_w", wt, u, v, w, g.dzhi, g.dx, g.dy, rhoref, rhorefh)

    def advec_2_s(self, st, s, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_2_s", st, s, u, v, w, g.dzi, g.dx, g.dy, rhoref, rhorefh)

    def advec_2_cfl(self, u, v, w, dt):
        g = self.g
        return self._call("ref_advec_2_cfl", u, v, w, g.dzi, g.dx, g.dy, float(dt), restype=C.c_double)

    # --- diff_2
    def diff_2_c(self, at, a, visc):
        g = self.g; self._call("ref_diff_2_c", at, a, float(visc), g.dx, g.dy, g.dzi, g.dzhi)

    def diff_2_w(self, wt, w, visc):
        g = self.g; self._call("ref_diff_2_w", wt, w, float(visc), g.dx, g.dy, g.dzi, g.dzhi)

    # --- advec_4 / diff_4 (grid built with order=4)
    def advec_4_u(self, ut, u, v, w):
        g = self.g; self._call("ref_advec_4_u", ut, u, v, w, g.dzi4, g.dx, g.dy)

    def advec_4_v(self, vt, u, v, w):
        g = self.g; self._call("ref_advec_4_v", vt, u, v, w, g.dzi4, g.dx, g.dy)

    def advec_4_w(self, wt, u, v, w):
        g = self.g; self._call("ref_advec_4_w", wt, u, v, w, g.dzhi4, g.dx, g.dy)

    def advec_4_s(self, st, s, u, v, w):
        g = self.g; self._call("ref_advec_4_s", st, s, u, v, w, g.dzi4, g.dx, g.dy)

    def advec_4_cfl(self, u, v, w, dt):
        g = self.g
        return self._call("ref_advec_4_cfl", u, v, w, g.dzi, g.dx, g.dy, float(dt), restype=C.c_double)

    def diff_4_c(self, at, a, visc):
        g = self.g; self._call("ref_diff_4_c", at, a, float(visc), g.dx, g.dy, g.dzi4, g.dzhi4)

    def diff_4_w(self, wt, w, visc):
        g = self.g; self._call("ref_diff_4_w", wt, w, float(visc), g.dx, g.dy, g.dzi4, g.dzhi4)

    # --- diff_smag2
    def diff_strain2(self, strain2, u, v, w, ugradbot, vgradbot, surface):
        g = self.g; TF = g.TF
        self._call("ref_diff_strain2", strain2, u, v, w, ugradbot, vgradbot, g.z, g.dzi, g.dzhi,
                   TF(1./float(g.dx)), TF(1./float(g.dy)), int(surface))

    def diff_evisc(self, evisc, u, v, w, N2, bgradbot, z0m, cs, tPr, surface, mason=True):
        g = self.g
        self._call("ref_diff_evisc", evisc, u, v, w, N2, bgradbot, g.z, g.dz, g.dzi, z0m, g.dx, g.dy,
                   float(cs), float(tPr), int(surface), int(mason))

    def diff_evisc_neutral(self, evisc, u, v, w, z0m, cs, visc, surface, mason=True):
        g = self.g
        self._call("ref_diff_evisc_neutral", evisc, u, v, w, None, None, g.z, g.dz, g.dzhi, z0m, g.dx, g.dy, g.zsize,
                   float(cs), float(visc), int(surface), int(mason))

    def diff_u(self, ut, u, v, w, evisc, fluxbot, fluxtop, rhoref, rhorefh, visc, surface):
        g = self.g; TF = g.TF
        self._call("ref_diff_u", ut, u, v, w, g.dzi, g.dzhi, TF(1./float(g.dx)), TF(1./float(g.dy)), evisc,
                   fluxbot, fluxtop, rhoref, rhorefh, float(visc), int(surface))

    def diff_v(self, vt, u, v, w, evisc, fluxbot, fluxtop, rhoref, rhorefh, visc, surface):
        g = self.g; TF = g.TF
        self._call("ref_diff_v", vt, u, v, w, g.dzi, g.dzhi, TF(1./float(g.dx)), TF(1./float(g.dy)), evisc,
                   fluxbot, fluxtop, rhoref, rhorefh, float(visc), int(surface))

    def diff_w(self, wt, u, v, w, evisc, rhoref, rhorefh, visc):
        g = self.g; TF = g.TF
        self._call("ref_diff_w", wt, u, v, w, g.dzi, g.dzhi, TF(1./float(g.dx)), TF(1./float(g.dy)), evisc,
                   rhoref, rhorefh, float(visc))

    def diff_c(self, at, a, evisc, fluxbot, fluxtop, rhoref, rhorefh, tPr, visc, surface):
        g = self.g; TF = g.TF
        self._call("ref_diff_c", at, a, g.dzi, g.dzhi, TF(1./(float(g.dx)*float(g.dx))), TF(1./(float(g.dy)*float(g.dy))),
                   evisc, fluxbot, fluxtop, rhoref, rhorefh, float(tPr), float(visc), int(surface))

    def diff_dnmul(self, evisc, tPr):
        g = self.g; TF = g.TF
        return self._call("ref_diff_dnmul", evisc, g.dzi, TF(1./(float(g.dx)*float(g.dx))), TF(1./(float(g.dy)*float(g.dy))),
                          float(tPr), restype=C.c_double)

    # --- thermo_dry
    def thermo_dry_N2(self, N2, th, thref):
        self._call("ref_thermo_dry_N2", N2, th, self.g.dzi, thref)

    def thermo_dry_buoyancy_tend_2nd(self, wt, th, threfh):
        self._call("ref_thermo_dry_buoyancy_tend_2nd", wt, th, threfh)

    # --- pres_2 tdma / rk3
    def tdma(self, a, b, c, p):
        kmax, jblock, iblock = p.shape
        work2d = np.zeros((jblock, iblock), self.g.TF)
        work3d = np.zeros_like(p)
        self._call("ref_pres_2_tdma", a, b, c, p, work2d, work3d, iblock, jblock, kmax)

    def rk3(self, a, at, substep, dt):
        self._call("ref_rk3", a, at, int(substep), float(dt))
