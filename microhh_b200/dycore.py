"""
Host-side mirror of the reference's operator interface for the hot path, over the C ABI.
Names follow the reference: `Fields` (maps mp/mt/sp/st/sd of include/fields.h:134-143),
`Boundary_cyclic.exec`, `Advec.exec/get_cfl`, `Diff.exec_viscosity/exec/get_dn`,
`Pres.exec/check_divergence`, `Timeloop.exec` (include/advec.h:54-60, include/diff.h:45-62,
include/pres.h:48-58, include/boundary_cyclic.h:42-50, include/timeloop.h:65).
torch is used for device memory only.
"""
import ctypes as C
import numpy as np
import torch

from . import capi
from .capi import FieldsC, ParamsC, GridDesc, SurfaceC, ForcingC, MhhError

SWADVEC = {"2i5": 25, "2": 2, "2i4": 24, "2i62": 262, "4": 4, "4m": 41}
SWDIFF = {"smag2": 1, "2": 2, "tke2": 3, "4": 4}
SWTHERMO = {"0": 0, None: 0, "disabled": 0, "dry": 1, "buoy": 2, "moist": 3}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class Context:
    """One mhh_ctx: grid + base state on one GPU."""

    def __init__(self, gd, device=0):
        if not torch.cuda.is_available():
            raise MhhError("mhhb200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = capi.load()
        self.gd = gd
        self.device = device
        self.torch_dtype = torch.float64 if gd.dtype == np.float64 else torch.float32
        self._keep = [np.ascontiguousarray(getattr(gd, n)) for n in ("z", "zh", "dz", "dzh", "dzi", "dzhi")]
        self._keep4 = [np.ascontiguousarray(getattr(gd, n)) for n in ("dzi4", "dzhi4")] if getattr(gd, "order", 2) == 4 else [None, None]
        d = GridDesc(gd.itot, gd.jtot, gd.ktot, gd.imax, gd.jmax, gd.kmax, gd.igc, gd.jgc, gd.kgc,
                     float(gd.xsize), float(gd.ysize), float(gd.zsize),
                     *[a.ctypes.data_as(C.c_void_p) for a in self._keep],
                     gd.npx, gd.npy, gd.mpicoordx, gd.mpicoordy,
                     *[None if a is None else a.ctypes.data_as(C.c_void_p) for a in self._keep4])
        h = C.c_void_p()
        rc = self.lib.mhh_ctx_create(C.byref(d), capi.MHH_F64 if gd.dtype == np.float64 else capi.MHH_F32,
                                     device, C.byref(h))
        self.h = h
        if rc != 0:
            msg = self.lib.mhh_last_error(h).decode()
            if h:
                self.lib.mhh_ctx_destroy(h)
            self.h = None
            raise MhhError(f"mhh_ctx_create failed ({rc}): {msg}")
        # run on torch's current stream so torch.cuda.Event timing sees the kernels
        self.use_torch_stream()
        if gd.npy > 1:
            self.comm_init()

    def comm_init(self):
        """Connect the y-slab ranks (one process per GPU): rank 0 makes the NCCL unique id, torch.distributed
        broadcasts the bytes (MicroHH's Master would use MPI_Bcast), every rank joins."""
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size() != self.gd.npy or dist.get_rank() != self.gd.mpicoordy:
            raise MhhError("slab context needs torch.distributed initialised with world_size = npy and rank = mpicoordy")
        n = capi.MHH_COMM_ID_BYTES
        buf = (C.c_ubyte * n)()
        if dist.get_rank() == 0:
            self.check(self.lib.mhh_comm_get_unique_id(buf, n))
        dev = torch.device("cuda", self.device) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0)
        buf = (C.c_ubyte * n)(*t.cpu().tolist())
        self.check(self.lib.mhh_comm_init(self.h, buf, n))
        # fused transposes: exchange the CUDA IPC handles of the spectral workspaces and map the peers' buffers
        if dist.get_backend() == "nccl":
            m = capi.MHH_IPC_BYTES
            hb = (C.c_ubyte * m)()
            self.check(self.lib.mhh_comm_get_ipc_handles(self.h, hb, m))
            mine = torch.tensor(list(hb), dtype=torch.uint8, device=dev)
            allh = [torch.empty_like(mine) for _ in range(self.gd.npy)]
            dist.all_gather(allh, mine)
            flat = torch.cat(allh).cpu().tolist()
            rc = self.lib.mhh_comm_open_peers(self.h, (C.c_ubyte * len(flat))(*flat), len(flat))
            # the transport is a collective property: if the mapping failed anywhere, every rank goes back to NCCL
            ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                import warnings
                warnings.warn("mhhb200: CUDA IPC peer mapping unavailable (%s); using NCCL transposes and halos"
                              % self.lib.mhh_last_error(self.h).decode())
                self.check(self.lib.mhh_comm_disable_peers(self.h))

    def use_torch_stream(self):
        s = torch.cuda.current_stream(self.device)
        self.check(self.lib.mhh_set_stream(self.h, C.c_void_p(s.cuda_stream or None)))

    def check(self, rc):
        if rc != 0:
            raise MhhError(f"mhhb200 error {rc}: {self.lib.mhh_last_error(self.h).decode()}")

    def set_basestate(self, rhoref, rhorefh, thref=None, threfh=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=self.gd.dtype) for a in (rhoref, rhorefh, thref, threfh)]
        self.check(self.lib.mhh_set_basestate(self.h, *[None if a is None else a.ctypes.data_as(C.c_void_p) for a in arrs]))

    def sync(self):
        self.check(self.lib.mhh_sync(self.h))

    @property
    def launch_count(self):
        return int(self.lib.mhh_launch_count(self.h))

    def profile_start(self):
        self.check(self.lib.mhh_profile_start(self.h))

    def profile_stop(self):
        import json
        out = C.c_char_p()
        self.check(self.lib.mhh_profile_stop(self.h, C.byref(out)))
        return json.loads(out.value.decode())

    @property
    def graph_replays(self):
        """steps of mhh_dycore_step / _step_host that ran as a replayed CUDA graph"""
        return int(self.lib.mhh_graph_replays(self.h))

    @property
    def transport(self):
        """'single', 'nccl' (grouped send/recv transposes and halos) or 'peer' (fused NVLink peer stores)"""
        return ("single", "nccl", "peer")[int(self.lib.mhh_comm_transport(self.h))]

    @property
    def workspace_bytes(self):
        return int(self.lib.mhh_workspace_bytes(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.mhh_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Fields:
    """Device fields in the reference's ghosted layout, as torch tensors (kcells, jcells, icells)."""

    def __init__(self, ctx, case=None, scalars=("th",), visc=1.e-5, svisc=1.e-5, fluxlimit_list=()):
        self.ctx = ctx
        self.fluxlimit_list = tuple(fluxlimit_list)        # [advec] fluxlimit_list of the .ini
        gd = ctx.gd
        self.scalars = list(scalars)
        dev = torch.device("cuda", ctx.device)
        z3 = lambda: torch.zeros(gd.shape, dtype=ctx.torch_dtype, device=dev)
        z2 = lambda: torch.zeros(gd.shape2d, dtype=ctx.torch_dtype, device=dev)
        self.t = {}
        for n in ["u", "v", "w", "ut", "vt", "wt", "evisc", "p"]:
            self.t[n] = z3()
        for s in self.scalars:
            self.t[s] = z3(); self.t[s + "t"] = z3()
        names2d = ["u_fluxbot", "u_fluxtop", "v_fluxbot", "v_fluxtop", "dudz_mo", "dvdz_mo", "dbdz_mo", "z0m",
                   "u_bot", "u_gradbot", "u_top", "u_gradtop", "v_bot", "v_gradbot", "v_top", "v_gradtop"]
        for s in self.scalars:
            names2d += [f"{s}_fluxbot", f"{s}_fluxtop", f"{s}_bot", f"{s}_gradbot", f"{s}_top", f"{s}_gradtop"]
        for n in names2d:
            self.t[n] = z2()
        self.visc = visc
        self.svisc = svisc
        if case is not None:
            self.upload(case)
        self._build()

    def upload(self, case):
        for n, t in self.t.items():
            if n in case and isinstance(case[n], np.ndarray):
                t.copy_(torch.from_numpy(np.ascontiguousarray(case[n])))

    def download(self, names=None):
        names = names or list(self.t.keys())
        return {n: self.t[n].cpu().numpy() for n in names}

    def __getitem__(self, n):
        return self.t[n]

    def _build(self):
        c = FieldsC()
        for n in ["u", "v", "w", "ut", "vt", "wt", "evisc", "p", "u_fluxbot", "u_fluxtop", "v_fluxbot", "v_fluxtop",
                  "dudz_mo", "dvdz_mo", "dbdz_mo", "z0m", "u_bot", "u_gradbot", "u_top", "u_gradtop",
                  "v_bot", "v_gradbot", "v_top", "v_gradtop"]:
            setattr(c, n, _ptr(self.t[n]))
        c.ns = len(self.scalars)
        c.visc = self.visc
        for i, s in enumerate(self.scalars):
            c.s[i] = self.t[s].data_ptr(); c.st[i] = self.t[s + "t"].data_ptr()
            c.svisc[i] = self.svisc
            c.s_fluxlimit[i] = 1 if s in self.fluxlimit_list else 0
            c.s_fluxbot[i] = self.t[f"{s}_fluxbot"].data_ptr(); c.s_fluxtop[i] = self.t[f"{s}_fluxtop"].data_ptr()
            c.s_bot[i] = self.t[f"{s}_bot"].data_ptr(); c.s_gradbot[i] = self.t[f"{s}_gradbot"].data_ptr()
            c.s_top[i] = self.t[f"{s}_top"].data_ptr(); c.s_gradtop[i] = self.t[f"{s}_gradtop"].data_ptr()
        self.c = c


def make_params(swadvec="2i5", swdiff="smag2", swthermo="dry", surface_model=True, sw_mason=True,
                cs=0.23, tPr=1./3., mbcbot=capi.BC_NEUMANN, mbctop=capi.BC_NEUMANN,
                sbcbot=capi.BC_NEUMANN, sbctop=capi.BC_NEUMANN, ns=1):
    p = ParamsC()
    p.swadvec = SWADVEC[swadvec]; p.swdiff = SWDIFF[swdiff]
    p.swthermo = SWTHERMO[swthermo]
    p.surface_model = int(surface_model); p.sw_mason = int(sw_mason)
    p.cs = cs; p.tPr = tPr
    p.mbcbot = mbcbot; p.mbctop = mbctop
    for i in range(capi.MHH_MAX_SCALARS):
        p.sbcbot[i] = sbcbot; p.sbctop[i] = sbctop
    return p


class Boundary_cyclic:
    def __init__(self, ctx):
        self.ctx = ctx

    def exec(self, fld, edge=capi.EDGE_BOTH):
        self.ctx.check(self.ctx.lib.mhh_boundary_cyclic(self.ctx.h, _ptr(fld), edge))

    def exec_2d(self, fld):
        self.ctx.check(self.ctx.lib.mhh_boundary_cyclic_2d(self.ctx.h, _ptr(fld)))


class Boundary:
    def __init__(self, ctx):
        self.ctx = ctx

    def set_ghost_cells_field(self, fld, bcbot, bot, gradbot, bctop, top, gradtop):
        self.ctx.check(self.ctx.lib.mhh_boundary_ghost_cells_2nd(
            self.ctx.h, _ptr(fld), bcbot, _ptr(bot), _ptr(gradbot), bctop, _ptr(top), _ptr(gradtop)))


class Boundary_4th:
    """Boundary<TF>::set_ghost_cells / set_ghost_cells_w on a 4th-order grid (src/boundary.cxx:776-922)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def set_ghost_cells_field(self, fld, bcbot, bot, gradbot, bctop, top, gradtop):
        self.ctx.check(self.ctx.lib.mhh_boundary_ghost_cells_4th(
            self.ctx.h, _ptr(fld), bcbot, _ptr(bot), _ptr(gradbot), bctop, _ptr(top), _ptr(gradtop)))

    def set_ghost_cells_w(self, w, conservation):
        self.ctx.check(self.ctx.lib.mhh_boundary_ghost_cells_w_4th(self.ctx.h, _ptr(w), int(bool(conservation))))


class Advec:
    def __init__(self, ctx, swadvec="2i5"):
        self.ctx = ctx; self.sw = SWADVEC[swadvec]

    def exec(self, fields):
        self.ctx.check(self.ctx.lib.mhh_advec_exec(self.ctx.h, self.sw, C.byref(fields.c)))

    def get_cfl(self, fields, dt):
        out = C.c_double()
        self.ctx.check(self.ctx.lib.mhh_advec_get_cfl(self.ctx.h, self.sw, C.byref(fields.c), dt, C.byref(out)))
        return out.value


class Diff:
    def __init__(self, ctx, params):
        self.ctx = ctx; self.prm = params

    def exec_viscosity(self, fields, n2=None):
        self.ctx.check(self.ctx.lib.mhh_diff_smag2_exec_viscosity(self.ctx.h, C.byref(fields.c), C.byref(self.prm), _ptr(n2)))

    def exec(self, fields):
        self.ctx.check(self.ctx.lib.mhh_diff_smag2_exec(self.ctx.h, C.byref(fields.c), C.byref(self.prm)))

    def get_dn(self, fields, dt):
        out = C.c_double()
        self.ctx.check(self.ctx.lib.mhh_diff_smag2_get_dn(self.ctx.h, C.byref(fields.c), C.byref(self.prm), dt, C.byref(out)))
        return out.value


class Diff_tke2:
    """Diff_tke2<TF> (src/diff_tke2.cxx): Deardorff SGS-TKE closure.  `sgstke` is one of the prognostic scalars of `fields`;
    the eddy viscosity for heat / scalars (`eviscs`) is owned here, like fields.sd["eviscs"] in the reference."""
    DEFAULTS = dict(ap=1.5, cf=2.5, ce1=0.19, ce2=0.51, cm=0.12, ch1=1., ch2=2., cn=0.76)       # src/diff_tke2.cxx:525-532

    def __init__(self, ctx, params, fields, sgstke="sgstke", **constants):
        self.ctx = ctx; self.prm = params
        self.eviscs = torch.zeros_like(fields["evisc"]) if params.swthermo != 0 else None
        c = capi.Tke2C()
        c.isgstke = fields.scalars.index(sgstke)
        c.eviscs = _ptr(self.eviscs)
        for k, v in {**self.DEFAULTS, **constants}.items():
            setattr(c, k, v)
        self.c = c

    def create(self, fields):
        """cold start: limit the initial field at Constants::sgstke_min (src/diff_tke2.cxx:641-660)"""
        self.ctx.check(self.ctx.lib.mhh_diff_tke2_create(self.ctx.h, _ptr(fields[fields.scalars[self.c.isgstke]])))

    def exec_viscosity(self, fields, n2=None):
        self.ctx.check(self.ctx.lib.mhh_diff_tke2_exec_viscosity(self.ctx.h, C.byref(fields.c), C.byref(self.prm), C.byref(self.c), _ptr(n2)))

    def exec(self, fields):
        self.ctx.check(self.ctx.lib.mhh_diff_tke2_exec(self.ctx.h, C.byref(fields.c), C.byref(self.prm), C.byref(self.c)))

    def get_dn(self, fields, dt):
        out = C.c_double()
        self.ctx.check(self.ctx.lib.mhh_diff_tke2_get_dn(self.ctx.h, C.byref(fields.c), C.byref(self.prm), C.byref(self.c), dt, C.byref(out)))
        return out.value

    def register(self):
        """Run the closure inside the fused sub-steps of a Dycore with swdiff = "tke2" (mhh_dycore_set_tke2)."""
        self.ctx.check(self.ctx.lib.mhh_dycore_set_tke2(self.ctx.h, C.byref(self.c)))

    def unregister(self):
        self.ctx.check(self.ctx.lib.mhh_dycore_set_tke2(self.ctx.h, None))


class Limiter:
    """Limiter<TF>::exec on one field (src/limiter.cxx:35-59, 117-129)."""
    SGSTKE_MIN = 1.e-7

    def __init__(self, ctx):
        self.ctx = ctx

    def exec(self, at, a, min_value, sub_dt):
        self.ctx.check(self.ctx.lib.mhh_limiter_exec(self.ctx.h, _ptr(at), _ptr(a), min_value, sub_dt))


class Diff_2:
    """Diff_2<TF> (src/diff_2.cxx): constant-viscosity diffusion of u, v, w and all scalars."""

    def __init__(self, ctx):
        self.ctx = ctx

    def exec(self, fields):
        self.ctx.check(self.ctx.lib.mhh_diff_2_exec(self.ctx.h, C.byref(fields.c)))

    def get_dn(self, fields, dt):
        out = C.c_double()
        self.ctx.check(self.ctx.lib.mhh_diff_2_get_dn(self.ctx.h, C.byref(fields.c), dt, C.byref(out)))
        return out.value


class Diff_4:
    """Diff_4<TF> (src/diff_4.cxx): 4th-order constant-viscosity diffusion (needs a 4th-order grid)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def exec(self, fields):
        self.ctx.check(self.ctx.lib.mhh_diff_4_exec(self.ctx.h, C.byref(fields.c)))


class Thermo_dry:
    def __init__(self, ctx):
        self.ctx = ctx

    def exec(self, fields):
        self.ctx.check(self.ctx.lib.mhh_thermo_dry_exec(self.ctx.h, _ptr(fields["wt"]), _ptr(fields[fields.scalars[0]])))

    def get_thermo_field_N2(self, out, fields):
        self.ctx.check(self.ctx.lib.mhh_thermo_dry_n2(self.ctx.h, _ptr(out), _ptr(fields[fields.scalars[0]])))


class Thermo_buoy:
    """Thermo_buoy<TF> (src/thermo_buoy.cxx): scalar 0 of the fields is the buoyancy b.  alpha / n2: [thermo] alpha, N2
    (slope-enabled thermodynamics when either is non-zero), utrans: [grid] utrans, swbaroclinic / dbdy_ls: [thermo]."""

    def __init__(self, ctx, alpha=0., n2=0., utrans=0., swbaroclinic=False, dbdy_ls=0.):
        self.ctx = ctx
        self.c = capi.ThermoBuoyC(float(alpha), float(n2), float(utrans), int(bool(swbaroclinic)), float(dbdy_ls))

    def exec(self, fields):
        self.ctx.check(self.ctx.lib.mhh_thermo_buoy_exec(self.ctx.h, C.byref(fields.c), C.byref(self.c)))

    def get_thermo_field_N2(self, out, fields):
        self.ctx.check(self.ctx.lib.mhh_thermo_buoy_n2(self.ctx.h, _ptr(out), _ptr(fields[fields.scalars[0]]), self.c.n2))

    def register(self):
        """Run thermo.exec inside the fused sub-steps of a Dycore with swthermo = "buoy" (mhh_dycore_set_thermo_buoy)."""
        self.ctx.check(self.ctx.lib.mhh_dycore_set_thermo_buoy(self.ctx.h, C.byref(self.c)))

    def unregister(self):
        self.ctx.check(self.ctx.lib.mhh_dycore_set_thermo_buoy(self.ctx.h, None))


class Thermo_moist:
    """Thermo_moist<TF> (src/thermo_moist.cxx): prognostic thl and qt (scalars `thl`, `qt` of the fields), saturation adjustment,
    hydrostatic base state kept in the context (thvref / thvrefh share the context's thref / threfh)."""
    PROFILES = ("pref", "prefh", "rhoref", "rhorefh", "thvref", "thvrefh", "exnref", "exnrefh")

    def __init__(self, ctx, fields, pbot, swupdatebasestate=True, thl="thl", qt="qt"):
        self.ctx = ctx
        self.c = capi.ThermoMoistC(fields.scalars.index(thl), fields.scalars.index(qt), float(pbot), int(bool(swupdatebasestate)))

    def _np(self, a):
        return np.ascontiguousarray(np.asarray(a, self.ctx.gd.dtype)[:self.ctx.gd.kcells])

    def calc_base_state(self, thl0, qt0):
        """create_basestate step 4: calc_base_state from the (ghosted) reference profiles of thl and qt"""
        a, b = self._np(thl0), self._np(qt0)
        self.ctx.check(self.ctx.lib.mhh_thermo_moist_calc_base_state(self.ctx.h, a.ctypes.data, b.ctypes.data, self.c.pbot))

    def set_profiles(self, **prof):
        keep = [self._np(prof[n]) if prof.get(n) is not None else None for n in self.PROFILES]
        self.ctx.check(self.ctx.lib.mhh_thermo_moist_set_profiles(self.ctx.h, *[a.ctypes.data if a is not None else None for a in keep]))

    def get_profiles(self):
        out = {n: np.zeros(self.ctx.gd.kcells, self.ctx.gd.dtype) for n in self.PROFILES}
        self.ctx.check(self.ctx.lib.mhh_thermo_moist_get_profiles(self.ctx.h, *[out[n].ctypes.data for n in self.PROFILES]))
        return out

    def exec(self, fields):
        self.ctx.check(self.ctx.lib.mhh_thermo_moist_exec(self.ctx.h, C.byref(fields.c), C.byref(self.c)))

    def get_thermo_field(self, out, name, fields):
        which = {"b": capi.MOIST_B, "ql": capi.MOIST_QL, "N2": capi.MOIST_N2}[name]
        self.ctx.check(self.ctx.lib.mhh_thermo_moist_get_thermo_field(self.ctx.h, which, _ptr(out), C.byref(fields.c), C.byref(self.c)))

    def get_buoyancy_surf(self, b, bbot, fields):
        self.ctx.check(self.ctx.lib.mhh_thermo_moist_get_buoyancy_surf(self.ctx.h, _ptr(b), _ptr(bbot), C.byref(fields.c), C.byref(self.c)))

    def get_buoyancy_fluxbot(self, bfluxbot, fields):
        self.ctx.check(self.ctx.lib.mhh_thermo_moist_get_buoyancy_fluxbot(self.ctx.h, _ptr(bfluxbot), C.byref(fields.c), C.byref(self.c)))

    def nonconverged(self):
        n = C.c_longlong()
        self.ctx.check(self.ctx.lib.mhh_thermo_moist_nonconverged(self.ctx.h, C.byref(n)))
        return n.value

    def base_state_sweeps(self):
        n = C.c_int()
        self.ctx.check(self.ctx.lib.mhh_thermo_moist_base_state_sweeps(self.ctx.h, C.byref(n)))
        return n.value

    def register(self):
        """Run thermo.exec inside the fused sub-steps of a Dycore with swthermo = "moist" (mhh_dycore_set_thermo_moist)."""
        self.ctx.check(self.ctx.lib.mhh_dycore_set_thermo_moist(self.ctx.h, C.byref(self.c)))

    def unregister(self):
        self.ctx.check(self.ctx.lib.mhh_dycore_set_thermo_moist(self.ctx.h, None))


class Pres:
    def __init__(self, ctx, swpres=2):
        self.ctx = ctx; self.sw = swpres

    def exec(self, fields, sub_dt):
        self.ctx.check(self.ctx.lib.mhh_pres_exec(self.ctx.h, self.sw, C.byref(fields.c), sub_dt))

    def check_divergence(self, fields):
        out = C.c_double()
        self.ctx.check(self.ctx.lib.mhh_pres_check_divergence(self.ctx.h, self.sw, C.byref(fields.c), C.byref(out)))
        return out.value

    def fft_roundtrip(self, a_in, a_out, solve=False):
        self.ctx.check(self.ctx.lib.mhh_pres_fft_roundtrip(self.ctx.h, _ptr(a_in), _ptr(a_out), int(solve)))


class Forcing:
    """Buffer<TF> (src/buffer.cxx) and Force<TF> (src/force.cxx): damping layer, large-scale pressure force, large-scale sources
    and subsidence.  Profiles are host arrays of kcells entries (uploaded here) or None."""
    LSPRES = {None: 0, "0": 0, "uflux": 1, "dpdx": 2, "geo": 3}

    def __init__(self, ctx, fields, swbuffer=False, zstart=0., sigma=2., beta=2., bufferprofs=None,
                 swlspres=None, uflux=0., dpdx=0., fc=0., ug=None, vg=None, utrans=0., vtrans=0., ls=None, wls=None):
        self.ctx = ctx
        dev = fields["u"].device
        self._keep = []

        def up(a):
            if a is None:
                return None
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=ctx.gd.dtype)).to(dev)
            self._keep.append(t)
            return t.data_ptr()
        c = ForcingC()
        c.swbuffer = int(bool(swbuffer)); c.buffer_zstart = zstart; c.buffer_sigma = sigma; c.buffer_beta = beta
        bp = bufferprofs or {}
        c.bufferprof_u = up(bp.get("u")); c.bufferprof_v = up(bp.get("v")); c.bufferprof_w = up(bp.get("w"))
        for i, n in enumerate(fields.scalars):
            c.bufferprof_s[i] = up(bp.get(n))
            c.ls_s[i] = up((ls or {}).get(n))
        c.swlspres = self.LSPRES[swlspres]; c.uflux = uflux; c.dpdx = dpdx; c.fc = fc
        c.ug = up(ug); c.vg = up(vg); c.utrans = utrans; c.vtrans = vtrans
        c.wls = up(wls)
        self.c = c

    def exec_buffer(self, fields):
        self.ctx.check(self.ctx.lib.mhh_buffer_exec(self.ctx.h, C.byref(fields.c), C.byref(self.c)))

    def exec_force(self, fields, sub_dt):
        self.ctx.check(self.ctx.lib.mhh_force_exec(self.ctx.h, C.byref(fields.c), C.byref(self.c), sub_dt))

    def register(self):
        """Run buffer.exec + force.exec inside the fused sub-steps from now on (mhh_dycore_set_forcing)."""
        self.ctx.check(self.ctx.lib.mhh_dycore_set_forcing(self.ctx.h, C.byref(self.c)))

    def unregister(self):
        self.ctx.check(self.ctx.lib.mhh_dycore_set_forcing(self.ctx.h, None))


class Boundary_surface:
    """Boundary_surface<TF> (src/boundary_surface.cxx): Monin-Obukhov surface model with constant z0 and the lookup solver.
    Owns the 2-D state (ustar, obuk, nobuk, z0m, z0h) like the reference class does."""
    SBC_DIRICHLET, SBC_FLUX = 0, 2

    def __init__(self, ctx, fields, z0m=0.1, z0h=0.1, thermobc=2, sbcbot=None):
        self.ctx = ctx
        gd = ctx.gd
        dev = fields["u"].device
        z2 = lambda v, dt=ctx.torch_dtype: torch.full(gd.shape2d, v, dtype=dt, device=dev)
        self.ustar = z2(1.e-9); self.obuk = z2(1.e-9)          # Constants::dsmall (init_surface)
        self.nobuk = z2(0, torch.int32)
        self.z0m = z2(z0m); self.z0h = z2(z0h); self.dutot = z2(0.)
        ctx.check(ctx.lib.mhh_boundary_surface_init(ctx.h, float(z0m), float(z0h), capi.BC_DIRICHLET, int(thermobc)))
        c = SurfaceC()
        c.ustar = self.ustar.data_ptr(); c.obuk = self.obuk.data_ptr(); c.nobuk = self.nobuk.data_ptr()
        c.z0m = self.z0m.data_ptr(); c.z0h = self.z0h.data_ptr(); c.dutot = self.dutot.data_ptr()
        sb = list(sbcbot) if sbcbot is not None else [thermobc] + [self.SBC_FLUX]*(len(fields.scalars) - 1)
        for i in range(capi.MHH_MAX_SCALARS):
            c.sbcbot[i] = sb[i] if i < len(sb) else 1
        self.c = c

    def exec(self, fields, prm):
        self.ctx.check(self.ctx.lib.mhh_boundary_surface_exec(self.ctx.h, C.byref(fields.c), C.byref(prm), C.byref(self.c)))


class Field3d_io:
    """Field3d_io<TF>::save_field3d / load_field3d (src/field3d_io.cxx): restart IO of one device field in the reference's
    file layout (interior as raw TF, no header); on y slabs every rank handles its rows of the one file."""

    def __init__(self, ctx):
        self.ctx = ctx

    def save_field3d(self, fld, filename, offset=0., kstart=None, kend=None):
        gd = self.ctx.gd
        return self.ctx.lib.mhh_field3d_save(self.ctx.h, _ptr(fld), str(filename).encode(), offset,
                                             gd.kstart if kstart is None else kstart, gd.kend if kend is None else kend)

    def load_field3d(self, fld, filename, offset=0., kstart=None, kend=None):
        gd = self.ctx.gd
        return self.ctx.lib.mhh_field3d_load(self.ctx.h, _ptr(fld), str(filename).encode(), offset,
                                             gd.kstart if kstart is None else kstart, gd.kend if kend is None else kend)


class Timeloop:
    def __init__(self, ctx):
        self.ctx = ctx

    def exec(self, fields, substep, dt):
        names = ["u", "v", "w"] + fields.scalars
        for n in names:
            self.ctx.check(self.ctx.lib.mhh_timeloop_rk3(self.ctx.h, _ptr(fields[n]), _ptr(fields[n + "t"]), substep, dt))


class Dycore:
    """Fused sub-step / step drivers (Model::exec order restricted to the hot path)."""

    def __init__(self, ctx, params):
        self.ctx = ctx; self.prm = params

    def substep(self, fields, substep, dt):
        self.ctx.check(self.ctx.lib.mhh_dycore_substep(self.ctx.h, C.byref(fields.c), C.byref(self.prm), substep, dt))

    def substep_pre(self, fields):
        self.ctx.check(self.ctx.lib.mhh_dycore_substep_pre(self.ctx.h, C.byref(fields.c), C.byref(self.prm)))

    def set_ghost_cells(self, fields):
        self.ctx.check(self.ctx.lib.mhh_dycore_set_ghost_cells(self.ctx.h, C.byref(fields.c), C.byref(self.prm)))

    def tendencies(self, fields):
        self.ctx.check(self.ctx.lib.mhh_dycore_tendencies(self.ctx.h, C.byref(fields.c), C.byref(self.prm)))

    def substep_post(self, fields, substep, dt):
        self.ctx.check(self.ctx.lib.mhh_dycore_substep_post(self.ctx.h, C.byref(fields.c), C.byref(self.prm), substep, dt))

    def substep_surface(self, fields, surface, substep, dt):
        """Model::exec order with the surface model on the device (src/model.cxx:368-504)."""
        self.ctx.check(self.ctx.lib.mhh_dycore_substep_surface(self.ctx.h, C.byref(fields.c), C.byref(self.prm), C.byref(surface.c), substep, dt))

    def step_surface(self, fields, surface, dt):
        for ss in range(3):
            self.substep_surface(fields, surface, ss, dt)

    def step(self, fields, dt):
        self.ctx.check(self.ctx.lib.mhh_dycore_step(self.ctx.h, C.byref(fields.c), C.byref(self.prm), dt))

    def step_host(self, fields, dt, nsteps, h_u, h_v, h_w, h_s):
        arr = (C.c_void_p * len(h_s))(*[t.data_ptr() for t in h_s])
        self.ctx.check(self.ctx.lib.mhh_dycore_step_host(
            self.ctx.h, C.byref(fields.c), C.byref(self.prm), dt, nsteps,
            C.c_void_p(h_u.data_ptr()), C.c_void_p(h_v.data_ptr()), C.c_void_p(h_w.data_ptr()), arr))
