"""
TEST INFRASTRUCTURE ONLY -- one RK3 sub-step / step of the dynamical core in the reference's
call order (Model<TF>::exec, reference src/model.cxx:356-504, restricted to the hot path):

  boundary.set_prognostic_cyclic_bcs (:368) -> boundary.set_ghost_cells (:370) ->
  diff.exec_viscosity (:376) -> thermo.exec (:388) -> advec.exec (:410) -> diff.exec (:414) ->
  pres.exec (:435-437) -> timeloop.exec (:504)

`K` supplies the kernels: oracle.NumpyKernels (the numpy restatement) or refbind.RefKernels
(the reference's own compiled CPU kernels).  The member-function glue (Pres_2 input/solve/
output, FFT) is always oracle.Pres2.
"""
import numpy as np
from . import oracle as O


def default_params(ns=1):
    return dict(swadvec="2i5", swdiff="smag2", surface_model=True, sw_mason=True, cs=0.23, tPr=1./3., swthermo="dry",
                mbcbot=O.BC_NEUMANN, mbctop=O.BC_NEUMANN, sbcbot=O.BC_NEUMANN, sbctop=O.BC_NEUMANN,
                visc=1.e-5, svisc=1.e-5)


def dycore_substep(g, K, c, prm, substep, dt, pres=None, timers=None, surface_model=None, forcing=None):
    """c: dict of numpy arrays as made by microhh_b200.synthetic.make_case (modified in place)."""
    import time
    t0 = time.perf_counter()
    def lap(name):
        nonlocal t0
        if timers is not None:
            t1 = time.perf_counter(); timers[name] = timers.get(name, 0.) + (t1 - t0); t0 = t1
    scal = c["scalars"]
    rr, rh = c["rhoref"], c["rhorefh"]
    surface = prm["surface_model"]
    for n in ["u", "v", "w"] + scal:
        K.boundary_cyclic(c[n])
    for n in ["u", "v"]:
        K.ghost_cells_bot_2nd(c[n], prm["mbcbot"], c.get(n + "_bot"), c.get(n + "_gradbot"))
        K.ghost_cells_top_2nd(c[n], prm["mbctop"], c.get(n + "_top"), c.get(n + "_gradtop"))
    for s in scal:
        K.ghost_cells_bot_2nd(c[s], prm["sbcbot"], c.get(s + "_bot"), c.get(s + "_gradbot"))
        K.ghost_cells_top_2nd(c[s], prm["sbctop"], c.get(s + "_top"), c.get(s + "_gradtop"))
    lap("boundary")
    swadvec, swdiff = prm.get("swadvec", "2i5"), prm.get("swdiff", "smag2")
    # exec_viscosity (Diff_smag2 only; Diff_2::exec_viscosity is empty)
    if swdiff == "smag2":
        K.diff_strain2(c["evisc"], c["u"], c["v"], c["w"], c["dudz_mo"], c["dvdz_mo"], surface)
        if prm["swthermo"] in ("dry", "moist"):
            N2 = np.zeros_like(c["evisc"])
            if prm["swthermo"] == "moist":
                K.thermo_moist_N2(N2, c["thl"], c["moist_bs"]["thvref"])       # Thermo_moist::get_thermo_field("N2"), src/thermo_moist.cxx:1612-1616
            else:
                K.thermo_dry_N2(N2, c[scal[0]], c["thref"])
            K.diff_evisc(c["evisc"], c["u"], c["v"], c["w"], N2, c["dbdz_mo"], c["z0m"], prm["cs"], prm["tPr"], surface, prm["sw_mason"])
        else:
            # Thermo_type::Disabled (src/diff_smag2.cxx:507-545)
            K.diff_evisc_neutral(c["evisc"], c["u"], c["v"], c["w"], c["z0m"], prm["cs"], prm["visc"], surface, prm["sw_mason"])
    elif swdiff == "tke2":
        _tke2_exec_viscosity(K, c, prm)
    lap("evisc")
    # thermo.exec
    if prm["swthermo"] == "dry":
        K.thermo_dry_buoyancy_tend_2nd(c["wt"], c[scal[0]], c["threfh"])
    elif prm["swthermo"] == "buoy":
        O.thermo_buoy_exec(K, c, prm.get("thermo_buoy", {}), 2)                 # Thermo_buoy::exec, Grid_order::Second branch
    elif prm["swthermo"] == "moist":
        thermo_moist_exec(K, c, prm["thermo_moist"])
    # boundary.exec (Monin-Obukhov surface model) + boundary.set_ghost_cells again (src/model.cxx:398-401)
    if surface_model is not None:
        surface_model.exec(c, c["thref"], c["threfh"], neutral=prm["swthermo"] != "dry")
        for n in ["u", "v"]:
            K.ghost_cells_bot_2nd(c[n], prm["mbcbot"], c.get(n + "_bot"), c.get(n + "_gradbot"))
            K.ghost_cells_top_2nd(c[n], prm["mbctop"], c.get(n + "_top"), c.get(n + "_gradtop"))
        for s_ in scal:
            K.ghost_cells_bot_2nd(c[s_], prm["sbcbot"], c.get(s_ + "_bot"), c.get(s_ + "_gradbot"))
            K.ghost_cells_top_2nd(c[s_], prm["sbctop"], c.get(s_ + "_top"), c.get(s_ + "_gradtop"))
    # advec.exec (Advec_2i5::exec src/advec_2i5.cxx:1017-1063, Advec_2::exec src/advec_2.cxx:311-345)
    A = {"2i5": (K.advec_2i5_u, K.advec_2i5_v, K.advec_2i5_w, K.advec_2i5_s),
         "2": (getattr(K, "advec_2_u", None), getattr(K, "advec_2_v", None), getattr(K, "advec_2_w", None), getattr(K, "advec_2_s", None)),
         # Advec_2i4::exec src/advec_2i4.cxx:700-738, Advec_2i62::exec src/advec_2i62.cxx:425-480
         "2i4": tuple(getattr(K, "advec_2i4_" + x, None) for x in "uvws"),
         "2i62": tuple(getattr(K, "advec_2i62_" + x, None) for x in "uvws")}[swadvec]
    A[0](c["ut"], c["u"], c["v"], c["w"], rr, rh)
    A[1](c["vt"], c["u"], c["v"], c["w"], rr, rh)
    A[2](c["wt"], c["u"], c["v"], c["w"], rr, rh)
    for s in scal:
        if swadvec in ("2i5", "2i62") and s in prm.get("fluxlimit_list", ()):
            K.advec_s_lim(c[s + "t"], c[s], c["u"], c["v"], c["w"], rr, rh)      # src/advec_2i5.cxx:1046-1056
        else:
            A[3](c[s + "t"], c[s], c["u"], c["v"], c["w"], rr, rh)
    lap("advec")
    # diff.exec
    if swdiff == "2":
        # Diff_2::exec (src/diff_2.cxx:163-190)
        K.diff_2_c(c["ut"], c["u"], prm["visc"])
        K.diff_2_c(c["vt"], c["v"], prm["visc"])
        K.diff_2_w(c["wt"], c["w"], prm["visc"])
        for s in scal:
            K.diff_2_c(c[s + "t"], c[s], prm["svisc"])
        lap("diff")
        if forcing is not None:
            forcing(c, O.rk3_subdt(dt, substep))       # buffer.exec + force.exec (src/model.cxx:416-430)
        return _pres_and_rk3(g, K, c, prm, substep, dt, pres, lap)
    if swdiff == "tke2":
        surface = True                                 # Diff_tke2::exec: Surface_model::Enabled throughout (src/diff_tke2.cxx:684-789)
    K.diff_u(c["ut"], c["u"], c["v"], c["w"], c["evisc"], c["u_fluxbot"], c["u_fluxtop"], rr, rh, prm["visc"], surface)
    K.diff_v(c["vt"], c["u"], c["v"], c["w"], c["evisc"], c["v_fluxbot"], c["v_fluxtop"], rr, rh, prm["visc"], surface)
    K.diff_w(c["wt"], c["u"], c["v"], c["w"], c["evisc"], rr, rh, prm["visc"])
    for s in scal:
        if swdiff == "tke2":
            # sgstke diffuses with the eddy viscosity for momentum, the other scalars with the one for heat when there is
            # buoyancy; tPr_dummy = 1 (src/diff_tke2.cxx:733-789)
            ev = c["evisc"] if (s == "sgstke" or prm["swthermo"] != "dry") else c["eviscs"]
            K.diff_c(c[s + "t"], c[s], ev, c[s + "_fluxbot"], c[s + "_fluxtop"], rr, rh, 1., prm["svisc"], True)
        else:
            K.diff_c(c[s + "t"], c[s], c["evisc"], c[s + "_fluxbot"], c[s + "_fluxtop"], rr, rh, prm["tPr"], prm["svisc"], surface)
    lap("diff")
    if forcing is not None:
        forcing(c, O.rk3_subdt(dt, substep))           # buffer.exec + force.exec (src/model.cxx:416-430)
    return _pres_and_rk3(g, K, c, prm, substep, dt, pres, lap)


def thermo_moist_exec(K, c, tm):
    """Thermo_moist::exec (src/thermo_moist.cxx:1415-1447); tm: dict(pbot, swupdatebasestate); c["moist_bs"]: the base state `bs`
    (updated in place when swupdatebasestate: Fields::exec's mean profiles, src/fields.cxx:542-551, then calc_base_state)."""
    if tm.get("swupdatebasestate", True):
        c["moist_bs"] = K.moist_base_state(K.mean_profile(c["thl"]), K.mean_profile(c["qt"]), tm["pbot"])
    bs = c["moist_bs"]
    K.thermo_moist_buoyancy_tend_2nd(c["wt"], c["thl"], c["qt"], bs["prefh"], bs["thvrefh"])


TKE2_DEFAULTS = dict(ap=1.5, cf=2.5, ce1=0.19, ce2=0.51, cm=0.12, ch1=1., ch2=2., cn=0.76)    # src/diff_tke2.cxx:525-532


def _tke2_exec_viscosity(K, c, prm):
    """Diff_tke2::exec_viscosity (src/diff_tke2.cxx:799-983): strain^2 with the MO gradients at the lowest level, eddy
    viscosities from the prognostic SGS TKE, and the buoyancy / dissipation / shear tendencies of sgstke."""
    t = {**TKE2_DEFAULTS, **prm.get("tke2", {})}
    mason = prm["sw_mason"]
    scal = c["scalars"]
    str2 = np.zeros_like(c["evisc"])
    K.diff_strain2(str2, c["u"], c["v"], c["w"], c["dudz_mo"], c["dvdz_mo"], True)
    if prm["swthermo"] != "dry":
        K.tke2_evisc_neutral(c["evisc"], c["sgstke"], c["u"], c["v"], c["w"], c["z0m"], t["cn"], t["cm"], mason)
        K.tke2_diss_tend_neutral(c["sgstket"], c["sgstke"], c["z0m"], t["ce1"], t["ce2"], mason)
    else:
        N2 = np.zeros_like(c["evisc"])
        K.thermo_dry_N2(N2, c[scal[0]], c["thref"])
        K.tke2_evisc(c["evisc"], c["sgstke"], c["u"], c["v"], c["w"], N2, c["dbdz_mo"], c["z0m"], t["cn"], t["cm"], mason)
        K.tke2_evisc_heat(c["eviscs"], c["evisc"], c["sgstke"], N2, c["dbdz_mo"], c["z0m"], t["cn"], t["ch1"], t["ch2"], mason)
        K.tke2_buoy_tend(c["sgstket"], c["sgstke"], c["eviscs"], N2, c["dbdz_mo"])
        K.tke2_diss_tend(c["sgstket"], c["sgstke"], N2, c["dbdz_mo"], c["z0m"], t["cn"], t["ce1"], t["ce2"], mason)
    K.tke2_shear_tend(c["sgstket"], c["sgstke"], c["evisc"], str2)


def _pres_and_rk3(g, K, c, prm, substep, dt, pres, lap):
    scal = c["scalars"]
    rr, rh = c["rhoref"], c["rhorefh"]
    # pres.exec(sub_dt)
    if pres is None:
        pres = O.Pres2(g, rr, rh)
    tdma = None
    if getattr(K, "tdma", None) is not None:
        tdma = lambda p, b: K.tdma(pres.a.copy(), b, pres.c.copy(), p)
    pres.exec(c["p"], c["u"], c["v"], c["w"], c["ut"], c["vt"], c["wt"], O.rk3_subdt(dt, substep), tdma)
    lap("pres")
    # limiter.exec: "apply the limiter as the last tendency" (src/model.cxx:439-440, src/limiter.cxx:117-129)
    if prm.get("swdiff") == "tke2":
        K.tendency_limiter(c["sgstket"], c["sgstke"], O.SGSTKE_MIN, O.rk3_subdt(dt, substep))
    # timeloop.exec
    for n in ["u", "v", "w"] + scal:
        K.rk3(c[n], c[n + "t"], substep, dt)
    lap("rk3")
    return pres


def dycore_substep_o4(g, K, c, prm, substep, dt, pres=None, forcing=None):
    """The 4th-order DNS configuration (swspatialorder=4: advec_4 or advec_4m + diff_4 + pres_4, no thermo), same call order of
    Model<TF>::exec (src/model.cxx:368-437): cyclic -> 4th-order ghost cells (w: normal type) -> w conservation type ->
    advec -> w normal -> diff -> w conservation -> pres -> w normal -> rk3."""
    scal = c["scalars"]
    for n in ["u", "v", "w"] + scal:
        K.boundary_cyclic(c[n])
    for n in ["u", "v"]:
        K.ghost_cells_bot_4th(c[n], prm["mbcbot"], c.get(n + "_bot"), c.get(n + "_gradbot"))
        K.ghost_cells_top_4th(c[n], prm["mbctop"], c.get(n + "_top"), c.get(n + "_gradtop"))
    K.ghost_cells_w_4th(c["w"], False)
    for s in scal:
        K.ghost_cells_bot_4th(c[s], prm["sbcbot"], c.get(s + "_bot"), c.get(s + "_gradbot"))
        K.ghost_cells_top_4th(c[s], prm["sbctop"], c.get(s + "_top"), c.get(s + "_gradtop"))
    if prm.get("swthermo") == "buoy":
        O.thermo_buoy_exec(K, c, prm.get("thermo_buoy", {}), 4)              # thermo.exec (src/model.cxx:388): w ghost cells still of the normal type
    K.ghost_cells_w_4th(c["w"], True)
    a4 = "advec_4m_" if prm.get("swadvec") == "4m" else "advec_4_"          # src/advec.cxx:79-84
    getattr(K, a4 + "u")(c["ut"], c["u"], c["v"], c["w"])
    getattr(K, a4 + "v")(c["vt"], c["u"], c["v"], c["w"])
    getattr(K, a4 + "w")(c["wt"], c["u"], c["v"], c["w"])
    for s in scal:
        getattr(K, a4 + "s")(c[s + "t"], c[s], c["u"], c["v"], c["w"])
    K.ghost_cells_w_4th(c["w"], False)
    K.diff_4_c(c["ut"], c["u"], prm["visc"])
    K.diff_4_c(c["vt"], c["v"], prm["visc"])
    K.diff_4_w(c["wt"], c["w"], prm["visc"])
    for s in scal:
        K.diff_4_c(c[s + "t"], c[s], prm["svisc"])
    if forcing is not None:
        forcing(c, O.rk3_subdt(dt, substep))       # buffer.exec + force.exec (src/model.cxx:416-430)
    K.ghost_cells_w_4th(c["w"], True)
    if pres is None:
        pres = O.Pres4(g)
    pres.exec(c["p"], c["u"], c["v"], c["w"], c["ut"], c["vt"], c["wt"], O.rk3_subdt(dt, substep))
    K.ghost_cells_w_4th(c["w"], False)
    for n in ["u", "v", "w"] + scal:
        K.rk3(c[n], c[n + "t"], substep, dt)
    return pres


def dycore_step(g, K, c, prm, dt, timers=None, surface_model=None, forcing=None, pres=None):
    """One RK3 step.  `pres`: the pressure solver object (default: the numpy Pres2 / Pres4 restatement; the tier-2 pin passes
    refbind.RefPres, the reference's own compiled Pres_2 / Pres_4 member functions)."""
    for ss in range(3):
        if prm.get("swadvec") in ("4", "4m"):
            pres = dycore_substep_o4(g, K, c, prm, ss, dt, pres, forcing)
        else:
            pres = dycore_substep(g, K, c, prm, ss, dt, pres, timers, surface_model, forcing)
    return pres
