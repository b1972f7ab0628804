"""
Generates tests/golden/new_*.npz: one full RK3 step of the rows added in the last session of round 2 (Thermo_moist, Thermo_buoy,
Advec_2i4, Advec_2i62) computed by the REFERENCE's own compiled CPU kernels and pressure solvers (oracle/_ref/libmhh_ref.so,
built from /root/reference by oracle/Makefile) in the reference's call order (oracle/step.py).  Run where /root/reference exists:

    make -C oracle && python tests/make_golden_new_rows.py

Inputs are regenerated from the case definition (tests/golden_new_rows.py, deterministic numpy generators); a checksum of the
inputs is stored so that a drift of the generators is detected rather than silently compared.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)

from golden_new_rows import CASES, NAMES, build, digest       # noqa: E402
from oracle import step as ostep, refbind                     # noqa: E402


def main():
    assert refbind.available(), "build oracle/_ref first (make -C oracle)"
    for name, (kind, dtype, o) in CASES.items():
        g, case, prm, dt, order = build(name)
        d = digest(case, NAMES[kind])
        R = refbind.RefKernels(g)
        pres = refbind.RefPres(g, order) if order == 4 else refbind.RefPres(g, 2, case["rhoref"], case["rhorefh"])
        ostep.dycore_step(g, R, case, prm, dt, pres=pres)
        out = {"step_" + n: case[n].copy() for n in NAMES[kind]}
        if kind == "moist":
            out.update({"bs_" + n: a.copy() for n, a in case["moist_bs"].items()})
        np.savez_compressed(os.path.join(HERE, "golden", "new_" + name + ".npz"), input_sha256=np.array(d), dt=np.array(dt), **out)
        print(name, d[:12], sorted(out)[:4])


if __name__ == "__main__":
    main()
