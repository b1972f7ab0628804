"""
Generates tests/golden/*.npz: outputs of the REFERENCE's own CPU kernels (oracle/_ref/libmhh_ref.so,
compiled from /root/reference by oracle/Makefile) for seeded synthetic inputs, in the reference's
call order (oracle/step.py).  Run in the build container where /root/reference exists:

    make -C oracle && python tests/make_golden.py

The inputs are not stored: they are regenerated from (shape, dtype, seed, flags) by
microhh_b200.synthetic.make_case, which is deterministic (numpy default_rng).  A checksum of the
inputs is stored so that a drift of the generator is detected rather than silently compared.
"""
import copy
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)

from util import make_pair, prepare_halos            # noqa: E402
from oracle import oracle as O, step as ostep, refbind  # noqa: E402

CASES = {
    # name: (shape, dtype, anelastic, stretched, nsteps)
    "step_16x12x8_f64": ((16, 12, 8), np.float64, False, False, 1),
    "step_20x12x10_anel_stretched_f64": ((20, 12, 10), np.float64, True, True, 1),
    "step_24x1x8_f64": ((24, 1, 8), np.float64, False, True, 1),
    "step_16x12x8_f32": ((16, 12, 8), np.float32, False, False, 1),
    "step_20x12x10_anel_stretched_f32": ((20, 12, 10), np.float32, True, True, 1),
}
DT = 2.0
VISC2, SVISC2 = 0.5, 0.7          # viscosities of the Diff_2 vectors (large enough to matter in one step)


def input_digest(case):
    h = hashlib.sha256()
    for n in ("u", "v", "w", "th", "rhoref", "rhorefh", "dudz_mo", "dbdz_mo", "th_fluxbot"):
        h.update(np.ascontiguousarray(case[n]).tobytes())
    return h.hexdigest()


# 4th-order DNS configuration (advec_4 + diff_4 + pres_4 + 4th-order ghost cells) on a 4th-order grid
CASES4 = {
    "step4_16x12x10_f64": ((16, 12, 10), np.float64, 0),
    "step4_24x1x8_f64": ((24, 1, 8), np.float64, 0),
    "step4_20x12x8_freeslip_f32": ((20, 12, 8), np.float32, 1),
}
VISC4, DT4 = 1e-3, 0.01


def make_case4(shape, dtype):
    """Inputs of the 4th-order vectors: a synthetic.make_case field set on a stretched 4th-order grid, w = 0 at the walls,
    a passive scalar, zero wall values / gradients."""
    from util import stretched_z
    from microhh_b200.grid import GridData
    from microhh_b200.synthetic import make_case
    it, jt, kt = shape
    z = stretched_z(kt, 2.)
    g = O.Grid(it, jt, kt, 6., 4., 2., 3, 3, 3, dtype, z=z, order=4)
    gd = GridData(it, jt, kt, 6., 4., 2., 3, 3, 3, dtype, z=z, order=4)
    case = make_case(gd, seed=5, noise=0.02)
    case["w"][:g.kstart+1] = 0; case["w"][g.kend:] = 0
    case["th"] = (1. + 0.1*case["u"]).astype(dtype)
    if jt == 1:
        case["v"][:] = 0
    for n in ("u", "v"):
        for sfx in ("_bot", "_top", "_gradbot", "_gradtop"):
            case[n + sfx] = np.zeros(gd.shape2d, dtype)
    case["th_gradbot"] = np.zeros(gd.shape2d, dtype); case["th_gradtop"] = np.zeros(gd.shape2d, dtype)
    return g, gd, case


def params4(mbc):
    prm = ostep.default_params()
    prm.update(swadvec="4", swdiff="4", visc=VISC4, svisc=VISC4, mbcbot=mbc, mbctop=mbc)
    return prm


def main4():
    for name, (shape, dtype, mbc) in CASES4.items():
        g, gd, case = make_case4(shape, dtype)
        digest = input_digest4(case)
        R = refbind.RefKernels(g)
        out = {}
        ck = copy.deepcopy(case)
        for n in ("u", "v", "w", "th"):
            R.boundary_cyclic(ck[n])
        R.advec_4_u(ck["ut"], ck["u"], ck["v"], ck["w"]); R.advec_4_v(ck["vt"], ck["u"], ck["v"], ck["w"])
        R.advec_4_w(ck["wt"], ck["u"], ck["v"], ck["w"]); R.advec_4_s(ck["tht"], ck["th"], ck["u"], ck["v"], ck["w"])
        for n in ("ut", "vt", "wt", "tht"):
            out["advec4_" + n] = ck[n].copy()
        out["cfl4"] = np.float64(R.advec_4_cfl(ck["u"], ck["v"], ck["w"], DT))
        R.diff_4_c(ck["ut"], ck["u"], VISC4); R.diff_4_c(ck["vt"], ck["v"], VISC4); R.diff_4_w(ck["wt"], ck["w"], VISC4)
        R.diff_4_c(ck["tht"], ck["th"], VISC4)
        for n in ("ut", "vt", "wt", "tht"):
            out["advdiff4_" + n] = ck[n].copy()
        cs = copy.deepcopy(case)
        ostep.dycore_step(g, refbind.RefKernels(g), cs, params4(mbc), DT4, pres=refbind.RefPres(g, 4))     # the reference's own Pres_4
        for n in ("u", "v", "w", "th", "p"):
            out["step_" + n] = cs[n].copy()
        np.savez_compressed(os.path.join(HERE, "golden", name + ".npz"), input_sha256=np.array(digest), shape=np.array(shape),
                            mbc=np.array(mbc), dt=np.array(DT4), **out)
        print(name, digest[:12])


def input_digest4(case):
    h = hashlib.sha256()
    for n in ("u", "v", "w", "th"):
        h.update(np.ascontiguousarray(case[n]).tobytes())
    return h.hexdigest()


def main():
    assert refbind.available(), "build oracle/_ref first (make -C oracle)"
    main4()
    for name, (shape, dtype, anel, stretched, nsteps) in CASES.items():
        g, gd, case = make_pair(*shape, dtype, stretched=stretched, anelastic=anel)
        digest = input_digest(case)
        out = {}
        # single kernels on halo-filled inputs (tendencies start at zero)
        ck = copy.deepcopy(case); prepare_halos(g, ck)
        R = refbind.RefKernels(g)
        rr, rh = ck["rhoref"], ck["rhorefh"]
        R.advec_2i5_u(ck["ut"], ck["u"], ck["v"], ck["w"], rr, rh)
        R.advec_2i5_v(ck["vt"], ck["u"], ck["v"], ck["w"], rr, rh)
        R.advec_2i5_w(ck["wt"], ck["u"], ck["v"], ck["w"], rr, rh)
        R.advec_2i5_s(ck["tht"], ck["th"], ck["u"], ck["v"], ck["w"], rr, rh)
        for n in ("ut", "vt", "wt", "tht"):
            out["advec_" + n] = ck[n].copy()
        out["cfl"] = np.float64(R.advec_2i5_cfl(ck["u"], ck["v"], ck["w"], DT))
        R.diff_strain2(ck["evisc"], ck["u"], ck["v"], ck["w"], ck["dudz_mo"], ck["dvdz_mo"], True)
        n2 = np.zeros_like(ck["evisc"]); R.thermo_dry_N2(n2, ck["th"], ck["thref"])
        R.diff_evisc(ck["evisc"], ck["u"], ck["v"], ck["w"], n2, ck["dbdz_mo"], ck["z0m"], 0.23, 1./3., True, True)
        out["evisc"] = ck["evisc"].copy()
        out["dn"] = np.float64(R.diff_dnmul(ck["evisc"], 1./3.)*DT)
        # Advec_2 / Diff_2 single kernels (tendencies start at zero again)
        c2 = copy.deepcopy(case); prepare_halos(g, c2)
        R.advec_2_u(c2["ut"], c2["u"], c2["v"], c2["w"], rr, rh)
        R.advec_2_v(c2["vt"], c2["u"], c2["v"], c2["w"], rr, rh)
        R.advec_2_w(c2["wt"], c2["u"], c2["v"], c2["w"], rr, rh)
        R.advec_2_s(c2["tht"], c2["th"], c2["u"], c2["v"], c2["w"], rr, rh)
        for n in ("ut", "vt", "wt", "tht"):
            out["advec2_" + n] = c2[n].copy()
        out["cfl2"] = np.float64(R.advec_2_cfl(c2["u"], c2["v"], c2["w"], DT))
        c3 = copy.deepcopy(case); prepare_halos(g, c3)
        R.diff_2_c(c3["ut"], c3["u"], VISC2); R.diff_2_c(c3["vt"], c3["v"], VISC2); R.diff_2_w(c3["wt"], c3["w"], VISC2)
        R.diff_2_c(c3["tht"], c3["th"], SVISC2)
        for n in ("ut", "vt", "wt", "tht"):
            out["diff2_" + n] = c3[n].copy()
        # full RK3 step with the plain 2nd-order schemes (swadvec=2, swdiff=2)
        c4 = copy.deepcopy(case)
        prm2 = ostep.default_params(); prm2.update(swadvec="2", swdiff="2", visc=VISC2, svisc=SVISC2)
        ostep.dycore_step(g, refbind.RefKernels(g), c4, prm2, DT, pres=refbind.RefPres(g, 2, c4["rhoref"], c4["rhorefh"]))
        for n in ("u", "v", "w", "th"):
            out["step22_" + n] = c4[n].copy()
        # full RK3 step(s) in Model::exec order
        cs = copy.deepcopy(case)
        prm = ostep.default_params()
        for _ in range(nsteps):
            ostep.dycore_step(g, refbind.RefKernels(g), cs, prm, DT, pres=refbind.RefPres(g, 2, cs["rhoref"], cs["rhorefh"]))     # the reference's own Pres_2
        for n in ("u", "v", "w", "th", "p"):
            out["step_" + n] = cs[n].copy()
        np.savez_compressed(os.path.join(HERE, "golden", name + ".npz"), input_sha256=np.array(digest),
                            shape=np.array(shape), anel=np.array(anel), stretched=np.array(stretched),
                            nsteps=np.array(nsteps), dt=np.array(DT), **out)
        print(name, digest[:12], {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(out.items())[:3]})


if __name__ == "__main__":
    main()
