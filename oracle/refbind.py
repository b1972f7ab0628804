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

    # --- advec_2i4 (oracle/ref/ref_advec_2i4.cpp)
    def advec_2i4_u(self, at, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_2i4_u", at, u, v, w, g.dzi, g.dx, g.dy, rhoref, rhorefh)

    def advec_2i4_v(self, at, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_2i4_v", at, u, v, w, g.dzi, g.dx, g.dy, rhoref, rhorefh)

    def advec_2i4_w(self, at, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_2i4_w", at, u, v, w, g.dzhi, g.dx, g.dy, rhoref, rhorefh)

    def advec_2i4_s(self, st, s, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_2i4_s", st, s, u, v, w, g.dzi, g.dx, g.dy, rhoref, rhorefh)

    def advec_2i4_cfl(self, u, v, w, dt):
        g = self.g
        return self._call("ref_advec_2i4_cfl", u, v, w, g.dzi, g.dx, g.dy, float(dt), restype=C.c_double)

    # --- advec_2i62 (oracle/ref/ref_advec_2i62.cpp)
    def advec_2i62_u(self, at, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_2i62_u", at, u, v, w, g.dzi, g.dx, g.dy, rhoref, rhorefh)

    def advec_2i62_v(self, at, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_2i62_v", at, u, v, w, g.dzi, g.dx, g.dy, rhoref, rhorefh)

    def advec_2i62_w(self, at, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_2i62_w", at, u, v, w, g.dzhi, g.dx, g.dy, rhoref, rhorefh)

    def advec_2i62_s(self, st, s, u, v, w, rhoref, rhorefh):
        g = self.g; self._call("ref_advec_2i62_s", st, s, u, v, w, g.dzi, g.dx, g.dy, rhoref, rhorefh)

    def advec_2i62_cfl(self, u, v, w, dt):
        g = self.g
        return self._call("ref_advec_2i62_cfl", u, v, w, g.dzi, g.dx, g.dy, float(dt), restype=C.c_double)

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

    def advec_4m_u(self, ut, u, v, w):
        g = self.g; self._call("ref_advec_4m_u", ut, u, v, w, g.dzi4, g.dx, g.dy)

    def advec_4m_v(self, vt, u, v, w):
        g = self.g; self._call("ref_advec_4m_v", vt, u, v, w, g.dzi4, g.dx, g.dy)

    def advec_4m_w(self, wt, u, v, w):
        g = self.g; self._call("ref_advec_4m_w", wt, u, v, w, g.dzhi4, g.dx, g.dy)

    def advec_4m_s(self, st, s, u, v, w):
        g = self.g; self._call("ref_advec_4m_s", st, s, u, v, w, g.dzi4, g.dx, g.dy)

    def advec_4m_cfl(self, u, v, w, dt):
        g = self.g
        return self._call("ref_advec_4m_cfl", u, v, w, g.dzi, g.dx, g.dy, float(dt), restype=C.c_double)

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

    # --- thermo_buoy (oracle/ref/ref_thermo_buoy.cpp)
    def thermo_buoy_N2(self, N2, b, bg_n2):
        self._call("ref_thermo_buoy_N2", N2, b, float(bg_n2), self.g.dzi)

    def thermo_buoy_tend(self, wt, b, order=2):
        self._call("ref_thermo_buoy_tend", wt, b, int(order))

    def thermo_buoy_tend_slope(self, ut, wt, bt, b, u, w, alpha, n2, utrans, order=2):
        self._call("ref_thermo_buoy_tend_slope", ut, wt, bt, b, u, w, float(alpha), float(n2), float(utrans), int(order))

    def thermo_buoy_baroclinic(self, bt, v, dbdy_ls, order=2):
        self._call("ref_thermo_buoy_baroclinic", bt, v, float(dbdy_ls), int(order))

    # --- thermo_moist (oracle/ref/ref_thermo_moist.cpp)
    def moist_base_state(self, thlmean, qtmean, pbot):
        g = self.g; TF = g.TF
        names = ("pref", "prefh", "rhoref", "rhorefh", "thvref", "thvrefh", "exnref", "exnrefh")
        out = {n: np.zeros(g.kcells, TF) for n in names}
        c = lambda a: np.ascontiguousarray(np.asarray(a, TF)[:g.kcells])
        self._call("ref_moist_base_state", *[out[n] for n in names], c(thlmean), c(qtmean), float(pbot), c(g.z), c(g.dz), c(g.dzh))
        return out

    def moist_top_and_bot(self, thl0, qt0):
        g = self.g; self._call("ref_moist_top_and_bot", thl0, qt0, g.z, g.zh, g.dzhi)

    def mean_profile(self, fld):
        g = self.g
        out = np.zeros(g.kcells, g.TF)
        self._call("ref_mean_profile", out, fld, int(g.itot), int(g.jtot))
        return out

    def thermo_moist_buoyancy_tend_2nd(self, wt, thl, qt, ph, thvrefh):
        self._call("ref_moist_buoyancy_tend_2nd", wt, thl, qt, ph, thvrefh)

    def thermo_moist_buoyancy(self, b, thl, qt, p, thvref):
        self._call("ref_moist_buoyancy", b, thl, qt, p, thvref)

    def thermo_moist_liquid_water(self, ql, thl, qt, p):
        self._call("ref_moist_liquid_water", ql, thl, qt, p)

    def thermo_moist_N2(self, N2, thl, thvref):
        self._call("ref_moist_N2", N2, thl, self.g.dzi, thvref)

    def thermo_moist_buoyancy_bot(self, b, bbot, thl, thlbot, qt, qtbot, thvref, thvrefh):
        self._call("ref_moist_buoyancy_bot", b, bbot, thl, thlbot, qt, qtbot, thvref, thvrefh)

    def thermo_moist_buoyancy_fluxbot(self, bfluxbot, thl, thlfluxbot, qt, qtfluxbot, thvrefh):
        self._call("ref_moist_buoyancy_fluxbot", bfluxbot, thl, thlfluxbot, qt, qtfluxbot, thvrefh)

    # --- diff_tke2 + limiter (oracle/ref/ref_diff_tke2.cpp)
    def tke2_enforce_min(self, sgstke):
        self._call("ref_tke2_enforce_min", sgstke)

    def tke2_evisc_neutral(self, evisc, sgstke, u, v, w, z0m, cn, cm, mason=True):
        g = self.g
        self._call("ref_tke2_evisc_neutral", evisc, sgstke, u, v, w, g.z, g.dz, z0m, g.dx, g.dy, float(cn), float(cm), int(mason))

    def tke2_evisc(self, evisc, sgstke, u, v, w, N2, bgradbot, z0m, cn, cm, mason=True):
        g = self.g
        self._call("ref_tke2_evisc", evisc, sgstke, u, v, w, N2, bgradbot, g.z, g.dz, z0m, g.dx, g.dy, float(cn), float(cm), int(mason))

    def tke2_evisc_heat(self, evisch, evisc, sgstke, N2, bgradbot, z0m, cn, ch1, ch2, mason=True):
        g = self.g
        self._call("ref_tke2_evisc_heat", evisch, evisc, sgstke, N2, bgradbot, g.z, g.dz, z0m, g.dx, g.dy,
                   float(cn), float(ch1), float(ch2), int(mason))

    def tke2_shear_tend(self, at, a, evisc, strain2):
        self._call("ref_tke2_shear_tend", at, a, evisc, strain2)

    def tke2_buoy_tend(self, at, a, evisch, N2, bgradbot):
        self._call("ref_tke2_buoy_tend", at, a, evisch, N2, bgradbot)

    def tke2_diss_tend(self, at, a, N2, bgradbot, z0m, cn, ce1, ce2, mason=True):
        g = self.g
        self._call("ref_tke2_diss_tend", at, a, N2, bgradbot, g.z, g.dz, z0m, g.dx, g.dy, float(cn), float(ce1), float(ce2), int(mason))

    def tke2_diss_tend_neutral(self, at, a, z0m, ce1, ce2, mason=True):
        g = self.g
        self._call("ref_tke2_diss_tend_neutral", at, a, g.z, g.dz, z0m, g.dx, g.dy, float(ce1), float(ce2), int(mason))

    def tendency_limiter(self, at, a, min_value, dt):
        self._call("ref_limiter", at, a, float(min_value), float(dt))

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


class RefSurface:
    """The reference's Monin-Obukhov surface solver (oracle/ref/ref_boundary_surface.cpp): same state and call signature as
    oracle.BoundarySurface, arithmetic by the reference's own compiled kernels."""

    def __init__(self, g, z0m, z0h, mbcbot, thermobc, fast=False):
        self.lib = C.CDLL(lib_path(fast))
        self.g = g
        TF = g.TF
        self.sfx = "f64" if TF == np.float64 else "f32"
        self.ct = C.c_double if TF == np.float64 else C.c_float
        self.lib.ref_set_geom(g.itot, g.jtot, g.ktot, g.igc, g.jgc, g.kgc)
        self.mbcbot, self.thermobc = mbcbot, thermobc
        self.z0m = np.full((g.jcells, g.icells), z0m, TF); self.z0h = np.full((g.jcells, g.icells), z0h, TF)
        self.ustar = np.full((g.jcells, g.icells), 1.e-9, TF); self.obuk = np.full((g.jcells, g.icells), 1.e-9, TF)
        n = getattr(self.lib, "ref_surface_nlut_" + self.sfx)()
        self.nobuk = np.zeros((g.jcells, g.icells), np.int32)
        self.zL_sl = np.zeros(n, np.float32); self.f_sl = np.zeros(n, np.float32)
        getattr(self.lib, "ref_surface_lut_" + self.sfx)(self._p(self.zL_sl), self._p(self.f_sl), self.ct(z0m), self.ct(z0h),
                                                       self.ct(float(g.z[g.kstart])), int(mbcbot), int(thermobc))

    @staticmethod
    def _p(a):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.c_void_p)

    def exec(self, c, thref=None, threfh=None, neutral=False):
        g = self.g; TF = g.TF
        self.lib.ref_set_geom(g.itot, g.jtot, g.ktot, g.igc, g.jgc, g.kgc)
        k = g.kstart
        dutot = np.zeros((g.jcells, g.icells), TF)
        if neutral:
            bfluxbot = np.zeros((g.jcells, g.icells), TF); b = np.zeros(g.shape if hasattr(g, "shape") else c["u"].shape, TF)
            bbot = np.zeros((g.jcells, g.icells), TF); db_ref = 0.
        else:
            name = c["scalars"][0]
            GRAV = TF(9.81)
            bbot = np.ascontiguousarray(GRAV/threfh[k]*(c[name + "_bot"] - threfh[k])).astype(TF)
            b = np.zeros_like(c[name]); b[k] = GRAV/thref[k]*(c[name][k] - thref[k])
            bfluxbot = np.ascontiguousarray(GRAV/threfh[k]*c[name + "_fluxbot"]).astype(TF)
            db_ref = float(GRAV/thref[k]*(thref[k] - threfh[k]))
        f = getattr(self.lib, "ref_surface_exec_" + self.sfx)
        P = self._p
        f(P(self.ustar), P(self.obuk), P(self.nobuk), P(dutot), P(c["u"]), P(c["v"]), P(c["u_bot"]), P(c["v_bot"]),
          P(c["u_fluxbot"]), P(c["v_fluxbot"]), P(c["u_gradbot"]), P(c["v_gradbot"]), P(bfluxbot), P(b), P(bbot), self.ct(db_ref),
          P(np.ascontiguousarray(g.z)), P(self.z0m), P(self.z0h), P(self.zL_sl), P(self.f_sl), int(self.mbcbot), int(self.thermobc),
          1, int(bool(neutral)), P(c["dudz_mo"]), P(c["dvdz_mo"]), P(c["dbdz_mo"]))
        fs = getattr(self.lib, "ref_surface_surfs_" + self.sfx)
        for name in c["scalars"]:
            bc = c.get(name + "_bcbot", self.thermobc if name == c["scalars"][0] else 2)
            fs(P(c[name + "_bot"]), P(c[name + "_gradbot"]), P(c[name + "_fluxbot"]), P(self.ustar), P(self.obuk), P(c[name]),
               P(self.z0h), self.ct(float(g.z[k])), int(bc))
        return dutot


class RefForcing:
    """Buffer / Force kernels of the reference (oracle/ref/ref_buffer_force.cpp), numpy arrays in place."""

    def __init__(self, g, fast=False):
        self.lib = C.CDLL(lib_path(fast))
        self.g = g
        self.sfx = "f64" if g.TF == np.float64 else "f32"
        self.ct = C.c_double if g.TF == np.float64 else C.c_float
        self.lib.ref_set_geom(g.itot, g.jtot, g.ktot, g.igc, g.jgc, g.kgc)

    def _f(self, name):
        return getattr(self.lib, f"{name}_{self.sfx}")

    @staticmethod
    def _p(a):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.c_void_p)

    def buffer(self, at, a, abuf, z, zstart, beta, sigma, bufferkstart):
        g = self.g; ct = self.ct
        self._f("ref_buffer")(self._p(at), self._p(a), self._p(abuf), self._p(np.ascontiguousarray(z)), ct(zstart), ct(float(g.zsize)),
                              ct(beta), ct(sigma), int(bufferkstart))

    def fixed_flux(self, ut, uflux, u_mean, ut_mean, utrans, dt):
        ct = self.ct
        self._f("ref_force_fixed_flux")(self._p(ut), ct(uflux), ct(float(u_mean)), ct(float(ut_mean)), ct(utrans), ct(dt))

    def coriolis(self, ut, vt, u, v, ug, vg, fc, ugrid, vgrid, order=2):
        ct = self.ct
        self._f("ref_force_coriolis")(self._p(ut), self._p(vt), self._p(u), self._p(v), self._p(ug), self._p(vg), ct(fc), ct(ugrid), ct(vgrid), int(order))

    def ls_source(self, st, sls):
        self._f("ref_force_ls_source")(self._p(st), self._p(sls))

    def wls_local(self, st, s, wls):
        self._f("ref_force_wls_local")(self._p(st), self._p(s), self._p(wls), self._p(np.ascontiguousarray(self.g.dzhi)))


# ---------------------------------------------------------------------------------------------------------------------
# Tier-2 pin of the pressure glue: the reference's own FFT<TF> (src/fft.cxx), Pres_2<TF> (src/pres_2.cxx) and Pres_4<TF>
# (src/pres_4.cxx) member functions, compiled where they lie and run on stand-in objects (oracle/ref/ref_fake_pres.h).
# FFTW is not in this image: the plans the reference creates forward their batched 1-D transform to `_fft_callback`
# below, i.e. to the oracle's own r2hc / hc2r -- so everything AROUND the 1-D transform (slice loops and strides, rhs,
# modified wave numbers, matrix build, tdma / hdma, ghost cells, the gradient) is the reference's compiled code.
# ---------------------------------------------------------------------------------------------------------------------
_FFT_CB_TYPE = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)


def _fft_callback(pin, pout, n, howmany, stride, dist, kind, is_float):
    from . import oracle as O
    dt = np.float32 if is_float else np.float64
    ct = C.c_float if is_float else C.c_double
    extent = (howmany - 1)*dist + (n - 1)*stride + 1
    isz = np.dtype(dt).itemsize
    def view(ptr):
        flat = np.ctypeslib.as_array((ct*extent).from_address(ptr))
        return np.lib.stride_tricks.as_strided(flat, shape=(howmany, n), strides=(dist*isz, stride*isz))
    x = np.ascontiguousarray(view(pin))
    view(pout)[...] = O.r2hc(x, 1) if kind == 0 else O.hc2r(x, 1)


_fft_cb_keepalive = _FFT_CB_TYPE(_fft_callback)


class RefPres:
    """Pres_2 (order 2) or Pres_4 (order 4) of the reference through its own init / set_values / input / solve / output."""

    def __init__(self, g, order, rhoref=None, rhorefh=None):
        self.lib = C.CDLL(lib_path(False))
        self.g = g; self.order = order
        TF = g.TF
        self.sfx = "f64" if TF == np.float64 else "f32"
        self.ct = C.c_double if TF == np.float64 else C.c_float
        self.lib.ref_set_fft_callback(_fft_cb_keepalive)
        arr = lambda a: np.ascontiguousarray(np.asarray(a, TF)[:g.kcells]) if a is not None else None
        self._z = arr(g.z)
        self._set_grid()
        for n in ("ref_fft_create", "ref_pres_2_create", "ref_pres_4_create"):
            self._f(n).restype = C.c_void_p
        self.fft = C.c_void_p(self._f("ref_fft_create")())
        if order == 2:
            rr, rh = arr(rhoref), arr(rhorefh)
            self.h = C.c_void_p(self._f("ref_pres_2_create")(self.fft, self._p(rr), self._p(rh), g.kcells))     # runs Pres_2::init -> FFT::init
        else:
            self.h = C.c_void_p(self._f("ref_pres_4_create")(self.fft, g.kcells))
        self._f("ref_fft_load")(self.fft)                                                                      # FFT::load: plan creation
        self.bmati = np.zeros(g.itot, TF); self.bmatj = np.zeros(g.jtot, TF)
        if order == 2:
            self.a = np.zeros(g.kmax, TF); self.c = np.zeros(g.kmax, TF)
            self._f("ref_pres_2_set_values")(self.h, self._p(self.bmati), self._p(self.bmatj), self._p(self.a), self._p(self.c))
        else:
            self.m = np.zeros((7, g.kmax), TF)
            self._f("ref_pres_4_set_values")(self.h, self._p(self.bmati), self._p(self.bmatj), self._p(self.m))

    def _set_grid(self):
        """(re)build the reference's own Grid<TF> for this case: Grid::init + Grid::calculate (oracle/ref/ref_grid.cpp); the
        Pres / FFT member functions then read the REFERENCE's metrics, not the oracle's"""
        g = self.g; ct = self.ct
        self._f("ref_grid_setup")(g.itot, g.jtot, g.ktot, ct(float(g.xsize)), ct(float(g.ysize)), ct(float(g.zsize)),
                                  g.igc, g.jgc, g.kgc, int(getattr(g, "order", 2)), self._p(self._z))

    def grid_metrics(self):
        """z, zh, dz, dzh, dzi, dzhi, dzi4, dzhi4 (kcells each) and (dx, dy, dxi, dyi, dzhi4bot, dzhi4top) of the reference's Grid"""
        self._set_grid()
        g = self.g
        out = np.zeros((8, g.kcells), g.TF); scal = np.zeros(6, g.TF)
        self._f("ref_grid_get")(self._p(out), self._p(scal))
        return dict(zip(("z", "zh", "dz", "dzh", "dzi", "dzhi", "dzi4", "dzhi4"), out)), scal

    def _f(self, name):
        return getattr(self.lib, f"{name}_{self.sfx}")

    @staticmethod
    def _p(a):
        if a is None:
            return None
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.c_void_p)

    def fft_forward(self, data):
        """FFT<TF>::exec_forward on a compact (kmax, jmax, imax) array, in place"""
        self._set_grid()
        tmp = np.zeros_like(data)
        self._f("ref_fft_forward")(self.fft, self._p(data), self._p(tmp))

    def fft_backward(self, data):
        """FFT<TF>::exec_backward: the result lands in the work array (src/fft.cxx:446-448); copied back here"""
        self._set_grid()
        tmp = np.zeros_like(data)
        self._f("ref_fft_backward")(self.fft, self._p(data), self._p(tmp))
        data[...] = tmp

    def input(self, p, u, v, w, ut, vt, wt, dt):
        self._set_grid()
        self._f(f"ref_pres_{self.order}_input")(self.h, self._p(p), self._p(u), self._p(v), self._p(w), self._p(ut), self._p(vt), self._p(wt), self.ct(dt))

    def solve(self, p):
        self._set_grid()
        tmp1 = np.zeros_like(p)
        if self.order == 2:
            tmp2 = np.zeros_like(p)
            self._f("ref_pres_2_solve")(self.h, self._p(p), self._p(tmp1), self._p(tmp2))
        else:
            self._f("ref_pres_4_solve")(self.h, self._p(p), self._p(tmp1))

    def output(self, ut, vt, wt, p):
        self._set_grid()
        self._f(f"ref_pres_{self.order}_output")(self.h, self._p(ut), self._p(vt), self._p(wt), self._p(p))

    def exec(self, p, u, v, w, ut, vt, wt, dt, tdma=None):
        """Pres_2::exec (src/pres_2.cxx:66-94) / Pres_4::exec (src/pres_4.cxx:77-144) without the statistics calls
        (`tdma` is accepted for call compatibility with oracle.Pres2.exec and ignored: the reference's own solver runs)"""
        p0 = p.copy()
        self.input(p, u, v, w, ut, vt, wt, dt)
        self.solve(p)
        self.output(ut, vt, wt, p)
        # The reference builds the compact rhs in place at the head of the p array; ghost locations that solve() does not
        # define afterwards (never read by anything) keep that scratch.  They are put back to their old contents here so
        # that whole-array comparisons against the oracle (which leaves them untouched) stay meaningful.
        defined = self.defined_mask()
        p[~defined] = p0[~defined]

    def defined_mask(self):
        """Where solve() leaves a defined pressure: the interior, the ghost levels it sets (one below for Pres_2, two below
        and above for Pres_4) over the interior columns, and what the cyclic fill derives from those."""
        g = self.g
        ks, ke = g.kstart, g.kend
        ng = 2 if self.order == 4 else 1
        nt = 2 if self.order == 4 else 0
        defined = np.zeros((g.kcells, g.jcells, g.icells), bool)
        if g.jtot > 1:                                # boundary_cyclic fills every level (src/boundary_cyclic.cxx:369-443) ...
            defined[ks-ng:ke+nt] = True
        else:                                         # ... but the y ghost rows only on the interior levels of a 2-D (jtot = 1) run
            defined[ks:ke] = True
            defined[ks-ng:ks, g.jstart:g.jend, :] = True
            defined[ke:ke+nt, g.jstart:g.jend, :] = True
        return defined

    def divergence(self, u, v, w):
        assert self.order == 4
        self._set_grid()
        f = self._f("ref_pres_4_divergence"); f.restype = C.c_double
        return f(self.h, self._p(u), self._p(v), self._p(w))


class RefField3dIO:
    """Field3d_io<TF>::save_field3d / load_field3d of the reference (oracle/ref/ref_field3d_io.cpp) on its own Grid<TF>."""

    def __init__(self, g):
        self.lib = C.CDLL(lib_path(False))
        self.g = g
        TF = g.TF
        self.sfx = "f64" if TF == np.float64 else "f32"
        self.ct = C.c_double if TF == np.float64 else C.c_float
        z = np.ascontiguousarray(np.asarray(g.z, TF)[:g.kcells])
        getattr(self.lib, f"ref_grid_setup_{self.sfx}")(
            g.itot, g.jtot, g.ktot, self.ct(float(g.xsize)), self.ct(float(g.ysize)), self.ct(float(g.zsize)),
            g.igc, g.jgc, g.kgc, int(getattr(g, "order", 2)), z.ctypes.data_as(C.c_void_p))

    def _io(self, name, data, filename, offset, kstart, kend):
        g = self.g
        assert data.flags["C_CONTIGUOUS"] and data.dtype == g.TF
        tmp1 = np.zeros_like(data); tmp2 = np.zeros_like(data)
        k0 = g.kstart if kstart is None else kstart
        k1 = g.kend if kend is None else kend
        f = getattr(self.lib, f"{name}_{self.sfx}"); f.restype = C.c_int
        return f(data.ctypes.data_as(C.c_void_p), tmp1.ctypes.data_as(C.c_void_p), tmp2.ctypes.data_as(C.c_void_p),
                 C.c_char_p(str(filename).encode()), self.ct(float(offset)), C.c_int(k0), C.c_int(k1))

    def save(self, data, filename, offset=0., kstart=None, kend=None):
        return self._io("ref_field3d_save", data, filename, offset, kstart, kend)

    def load(self, data, filename, offset=0., kstart=None, kend=None):
        return self._io("ref_field3d_load", data, filename, offset, kstart, kend)
