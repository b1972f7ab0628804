"""Golden vectors of Thermo_moist, Thermo_buoy, Advec_2i4 and Advec_2i62 (tests/golden/new_*.npz, made by the reference's own
compiled kernels, tests/make_golden_new_rows.py): the numpy oracle reproduces them bit for bit from the regenerated inputs.
Needs neither /root/reference nor a GPU, so the oracle stays pinned on the GPU box too."""
import glob
import os

import numpy as np
import pytest

from golden_new_rows import CASES, NAMES, build, digest
from oracle import oracle as O, step as ostep

HERE = os.path.dirname(os.path.abspath(__file__))


def test_golden_new_rows_present():
    have = {os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(HERE, "golden", "new_*.npz"))}
    assert have == set(CASES), (have, set(CASES))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden_new_rows_bitexact(name):
    z = np.load(os.path.join(HERE, "golden", "new_" + name + ".npz"))
    kind = CASES[name][0]
    g, case, prm, dt, order = build(name)
    assert digest(case, NAMES[kind]) == str(z["input_sha256"]), "the input generator drifted"
    assert dt == float(z["dt"])
    ostep.dycore_step(g, O.NumpyKernels(g), case, prm, dt)
    for n in NAMES[kind]:
        assert np.array_equal(case[n], z["step_" + n]), n
    if kind == "moist":
        for n, a in case["moist_bs"].items():
            assert np.array_equal(a, z["bs_" + n]), n
