"""
ctypes binding of libmhhb200.so (include/mhhb200.h).  This is the ONLY compute path of the
package: if the CUDA extension is missing or no CUDA device is present, it raises -- there
is no CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MHH_LIB: another build of the same library (A/B measurements of compile-time switches, tools/gpu_*.sh)
LIB_PATH = os.environ.get("MHH_LIB") or os.path.join(_HERE, "lib", "libmhhb200.so")

MHH_F64, MHH_F32 = 0, 1
MHH_MAX_SCALARS = 8
MHH_COMM_ID_BYTES = 128
MHH_IPC_BYTES = 192
EDGE_EAST_WEST, EDGE_NORTH_SOUTH, EDGE_BOTH = 0, 1, 2
BC_NONE, BC_DIRICHLET, BC_NEUMANN = -1, 0, 1

_vp = C.c_void_p
_SA = _vp * MHH_MAX_SCALARS


class GridDesc(C.Structure):
    _fields_ = [("itot", C.c_int), ("jtot", C.c_int), ("ktot", C.c_int),
                ("imax", C.c_int), ("jmax", C.c_int), ("kmax", C.c_int),
                ("igc", C.c_int), ("jgc", C.c_int), ("kgc", C.c_int),
                ("xsize", C.c_double), ("ysize", C.c_double), ("zsize", C.c_double),
                ("z", _vp), ("zh", _vp), ("dz", _vp), ("dzh", _vp), ("dzi", _vp), ("dzhi", _vp),
                ("npx", C.c_int), ("npy", C.c_int), ("mpicoordx", C.c_int), ("mpicoordy", C.c_int),
                ("dzi4", _vp), ("dzhi4", _vp)]


class FieldsC(C.Structure):
    _fields_ = [("u", _vp), ("v", _vp), ("w", _vp), ("ut", _vp), ("vt", _vp), ("wt", _vp),
                ("evisc", _vp), ("p", _vp), ("ns", C.c_int),
                ("s", _SA), ("st", _SA), ("svisc", C.c_double * MHH_MAX_SCALARS), ("visc", C.c_double),
                ("u_fluxbot", _vp), ("u_fluxtop", _vp), ("v_fluxbot", _vp), ("v_fluxtop", _vp),
                ("s_fluxbot", _SA), ("s_fluxtop", _SA),
                ("dudz_mo", _vp), ("dvdz_mo", _vp), ("dbdz_mo", _vp), ("z0m", _vp),
                ("u_bot", _vp), ("u_gradbot", _vp), ("u_top", _vp), ("u_gradtop", _vp),
                ("v_bot", _vp), ("v_gradbot", _vp), ("v_top", _vp), ("v_gradtop", _vp),
                ("s_bot", _SA), ("s_gradbot", _SA), ("s_top", _SA), ("s_gradtop", _SA),
                ("s_fluxlimit", C.c_int * MHH_MAX_SCALARS)]


class Slab2Info(C.Structure):
    _fields_ = [("nm", C.c_int), ("mcl", C.c_int), ("m_off", C.c_int), ("jmax", C.c_int), ("npan", C.c_int), ("ksplit", C.c_int),
                ("xside_elems", C.c_longlong), ("yside_elems", C.c_longlong)]


class ForcingC(C.Structure):
    _fields_ = [("swbuffer", C.c_int), ("buffer_zstart", C.c_double), ("buffer_sigma", C.c_double), ("buffer_beta", C.c_double),
                ("bufferprof_u", C.c_void_p), ("bufferprof_v", C.c_void_p), ("bufferprof_w", C.c_void_p), ("bufferprof_s", _SA),
                ("swlspres", C.c_int), ("uflux", C.c_double), ("dpdx", C.c_double), ("fc", C.c_double),
                ("ug", C.c_void_p), ("vg", C.c_void_p), ("utrans", C.c_double), ("vtrans", C.c_double),
                ("ls_s", _SA), ("wls", C.c_void_p)]


class Tke2C(C.Structure):
    _fields_ = [("isgstke", C.c_int), ("eviscs", C.c_void_p),
                ("ap", C.c_double), ("cf", C.c_double), ("ce1", C.c_double), ("ce2", C.c_double),
                ("cm", C.c_double), ("ch1", C.c_double), ("ch2", C.c_double), ("cn", C.c_double)]


class ThermoBuoyC(C.Structure):
    _fields_ = [("alpha", C.c_double), ("n2", C.c_double), ("utrans", C.c_double), ("swbaroclinic", C.c_int), ("dbdy_ls", C.c_double)]


class ThermoMoistC(C.Structure):
    _fields_ = [("ithl", C.c_int), ("iqt", C.c_int), ("pbot", C.c_double), ("swupdatebasestate", C.c_int)]


MOIST_B, MOIST_QL, MOIST_N2 = 0, 1, 2


class SurfaceC(C.Structure):
    _fields_ = [("ustar", C.c_void_p), ("obuk", C.c_void_p), ("nobuk", C.c_void_p), ("z0m", C.c_void_p), ("z0h", C.c_void_p),
                ("dutot", C.c_void_p), ("sbcbot", C.c_int * MHH_MAX_SCALARS)]


class SlabInfo(C.Structure):
    _fields_ = [("nm", C.c_int), ("mcl", C.c_int), ("m_off", C.c_int), ("jmax", C.c_int),
                ("rows", C.c_longlong), ("xside_elems", C.c_longlong), ("yside_elems", C.c_longlong)]


class ParamsC(C.Structure):
    _fields_ = [("swadvec", C.c_int), ("swdiff", C.c_int), ("swthermo", C.c_int),
                ("surface_model", C.c_int), ("sw_mason", C.c_int),
                ("cs", C.c_double), ("tPr", C.c_double),
                ("mbcbot", C.c_int), ("mbctop", C.c_int),
                ("sbcbot", C.c_int * MHH_MAX_SCALARS), ("sbctop", C.c_int * MHH_MAX_SCALARS)]


# name -> (restype, argtypes); must list every symbol declared in include/mhhb200.h
_PF, _PP = C.POINTER(FieldsC), C.POINTER(ParamsC)
SIGNATURES = {
    "mhh_ctx_create": (C.c_int, [C.POINTER(GridDesc), C.c_int, C.c_int, C.POINTER(_vp)]),
    "mhh_ctx_destroy": (None, [_vp]),
    "mhh_last_error": (C.c_char_p, [_vp]),
    "mhh_sync": (C.c_int, [_vp]),
    "mhh_set_stream": (C.c_int, [_vp, _vp]),
    "mhh_set_basestate": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "mhh_launch_count": (C.c_longlong, [_vp]),
    "mhh_workspace_bytes": (C.c_longlong, [_vp]),
    "mhh_profile_start": (C.c_int, [_vp]),
    "mhh_profile_stop": (C.c_int, [_vp, C.POINTER(C.c_char_p)]),
    "mhh_comm_get_unique_id": (C.c_int, [_vp, C.c_int]),
    "mhh_comm_init": (C.c_int, [_vp, _vp, C.c_int]),
    "mhh_comm_get_ipc_handles": (C.c_int, [_vp, _vp, C.c_int]),
    "mhh_comm_open_peers": (C.c_int, [_vp, _vp, C.c_int]),
    "mhh_comm_disable_peers": (C.c_int, [_vp]),
    "mhh_comm_transport": (C.c_int, [_vp]),
    "mhh_slab_layout": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(SlabInfo)]),
    "mhh_slab_xindex": (C.c_longlong, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int]),
    "mhh_slab_xindex_tiled": (C.c_longlong, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int, C.POINTER(C.c_longlong)]),
    "mhh_slab2_layout": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Slab2Info)]),
    "mhh_slab2_yindex": (C.c_longlong, [C.c_int]*9),
    "mhh_slab2_xindex": (C.c_longlong, [C.c_int]*8),
    "mhh_slab_yindex": (C.c_longlong, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mhh_boundary_cyclic": (C.c_int, [_vp, _vp, C.c_int]),
    "mhh_boundary_cyclic_2d": (C.c_int, [_vp, _vp]),
    "mhh_boundary_ghost_cells_2nd": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp]),
    "mhh_boundary_ghost_cells_4th": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp]),
    "mhh_boundary_ghost_cells_w_4th": (C.c_int, [_vp, _vp, C.c_int]),
    "mhh_advec_exec": (C.c_int, [_vp, C.c_int, _PF]),
    "mhh_advec_get_cfl": (C.c_int, [_vp, C.c_int, _PF, C.c_double, C.POINTER(C.c_double)]),
    "mhh_diff_smag2_exec_viscosity": (C.c_int, [_vp, _PF, _PP, _vp]),
    "mhh_diff_smag2_exec": (C.c_int, [_vp, _PF, _PP]),
    "mhh_diff_smag2_get_dn": (C.c_int, [_vp, _PF, _PP, C.c_double, C.POINTER(C.c_double)]),
    "mhh_diff_tke2_create": (C.c_int, [_vp, _vp]),
    "mhh_diff_tke2_exec_viscosity": (C.c_int, [_vp, _PF, _PP, C.POINTER(Tke2C), _vp]),
    "mhh_diff_tke2_exec": (C.c_int, [_vp, _PF, _PP, C.POINTER(Tke2C)]),
    "mhh_diff_tke2_get_dn": (C.c_int, [_vp, _PF, _PP, C.POINTER(Tke2C), C.c_double, C.POINTER(C.c_double)]),
    "mhh_limiter_exec": (C.c_int, [_vp, _vp, _vp, C.c_double, C.c_double]),
    "mhh_dycore_set_tke2": (C.c_int, [_vp, C.POINTER(Tke2C)]),
    "mhh_diff_2_exec": (C.c_int, [_vp, _PF]),
    "mhh_diff_4_exec": (C.c_int, [_vp, _PF]),
    "mhh_diff_2_get_dn": (C.c_int, [_vp, _PF, C.c_double, C.POINTER(C.c_double)]),
    "mhh_thermo_dry_exec": (C.c_int, [_vp, _vp, _vp]),
    "mhh_thermo_dry_n2": (C.c_int, [_vp, _vp, _vp]),
    "mhh_thermo_buoy_exec": (C.c_int, [_vp, _PF, C.POINTER(ThermoBuoyC)]),
    "mhh_thermo_buoy_n2": (C.c_int, [_vp, _vp, _vp, C.c_double]),
    "mhh_dycore_set_thermo_buoy": (C.c_int, [_vp, C.POINTER(ThermoBuoyC)]),
    "mhh_thermo_moist_calc_base_state": (C.c_int, [_vp, _vp, _vp, C.c_double]),
    "mhh_thermo_moist_set_profiles": (C.c_int, [_vp] + [_vp]*8),
    "mhh_thermo_moist_get_profiles": (C.c_int, [_vp] + [_vp]*8),
    "mhh_thermo_moist_exec": (C.c_int, [_vp, _PF, C.POINTER(ThermoMoistC)]),
    "mhh_thermo_moist_get_thermo_field": (C.c_int, [_vp, C.c_int, _vp, _PF, C.POINTER(ThermoMoistC)]),
    "mhh_thermo_moist_get_buoyancy_surf": (C.c_int, [_vp, _vp, _vp, _PF, C.POINTER(ThermoMoistC)]),
    "mhh_thermo_moist_get_buoyancy_fluxbot": (C.c_int, [_vp, _vp, _PF, C.POINTER(ThermoMoistC)]),
    "mhh_thermo_moist_nonconverged": (C.c_int, [_vp, C.POINTER(C.c_longlong)]),
    "mhh_thermo_moist_base_state_sweeps": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "mhh_dycore_set_thermo_moist": (C.c_int, [_vp, C.POINTER(ThermoMoistC)]),
    "mhh_pres_exec": (C.c_int, [_vp, C.c_int, _PF, C.c_double]),
    "mhh_pres_check_divergence": (C.c_int, [_vp, C.c_int, _PF, C.POINTER(C.c_double)]),
    "mhh_pres_fft_roundtrip": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "mhh_timeloop_rk3": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_double]),
    "mhh_dycore_substep": (C.c_int, [_vp, _PF, _PP, C.c_int, C.c_double]),
    "mhh_dycore_substep_pre": (C.c_int, [_vp, _PF, _PP]),
    "mhh_dycore_set_ghost_cells": (C.c_int, [_vp, _PF, _PP]),
    "mhh_dycore_tendencies": (C.c_int, [_vp, _PF, _PP]),
    "mhh_dycore_substep_post": (C.c_int, [_vp, _PF, _PP, C.c_int, C.c_double]),
    "mhh_buffer_exec": (C.c_int, [_vp, _PF, C.POINTER(ForcingC)]),
    "mhh_force_exec": (C.c_int, [_vp, _PF, C.POINTER(ForcingC), C.c_double]),
    "mhh_dycore_set_forcing": (C.c_int, [_vp, C.POINTER(ForcingC)]),
    "mhh_boundary_surface_init": (C.c_int, [_vp, C.c_double, C.c_double, C.c_int, C.c_int]),
    "mhh_boundary_surface_exec": (C.c_int, [_vp, _PF, _PP, C.POINTER(SurfaceC)]),
    "mhh_dycore_substep_surface": (C.c_int, [_vp, _PF, _PP, C.POINTER(SurfaceC), C.c_int, C.c_double]),
    "mhh_field3d_save": (C.c_int, [_vp, _vp, C.c_char_p, C.c_double, C.c_int, C.c_int]),
    "mhh_field3d_load": (C.c_int, [_vp, _vp, C.c_char_p, C.c_double, C.c_int, C.c_int]),
    "mhh_dycore_step": (C.c_int, [_vp, _PF, _PP, C.c_double]),
    "mhh_graph_replays": (C.c_longlong, [_vp]),
    "mhh_dycore_step_host": (C.c_int, [_vp, _PF, _PP, C.c_double, C.c_int, _vp, _vp, _vp, C.POINTER(_vp)]),
}

_lib = None


def load():
    """Load libmhhb200.so; raises if it has not been built (python __graft_entry__.py / make)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C microhh_b200/csrc` "
                "(or __graft_entry__.build()).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class MhhError(RuntimeError):
    pass
