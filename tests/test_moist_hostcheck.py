"""
The product's moist arithmetic without a GPU: microhh_b200/csrc/thermo_moist_kernels.cuh declares the saturation adjustment,
the buoyancy formulas and the serial base-state integration __host__ __device__; tests/hostcheck/moist_host_check.cu compiles
those very functions for the CPU (nvcc, host code only, no FMA contraction on x86-64) and this test holds them BIT FOR BIT
against the oracle, which is itself pinned bit for bit to the reference (tests/test_oracle_vs_ref.py).  What remains for the
`-m gpu` tests is the indexing of the kernels and the device's own exp / pow / FMA rounding.
"""
import ctypes as C
import os
import shutil
import subprocess
import numpy as np
import pytest

from util import make_pair, moist_case
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
DTYPES = [np.float64, np.float32]


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    out = str(tmp_path_factory.mktemp("hostcheck") / "libmoist_host_check.so")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
           "-Xcompiler", "-ffp-contract=off", "-shared", "-o", out, os.path.join(HERE, "hostcheck", "moist_host_check.cu")]
    subprocess.run(cmd, check=True)
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cold", [False, True])
def test_moist_arithmetic_host_bitexact(lib, dtype, cold):
    sfx = "f64" if dtype == np.float64 else "f32"
    ct = C.c_double if dtype == np.float64 else C.c_float
    g, gd, case = make_pair(16, 12, 24, dtype, stretched=True, sizes=(3200., 3200., 3000.))
    thl, qt = moist_case(g, gd, dtype, cold=cold)
    pbot = 70000. if cold else 101500.
    thl0, qt0 = O.mean_profile(g, thl), O.mean_profile(g, qt)
    O.moist_top_and_bot(g, thl0, qt0)
    ref = O.moist_base_state(g, thl0, qt0, pbot)
    names = ("pref", "prefh", "rhoref", "rhorefh", "thvref", "thvrefh", "exnref", "exnrefh")
    got = {n: np.zeros(g.kcells, dtype) for n in names}
    c1 = lambda a: np.ascontiguousarray(np.asarray(a, dtype)[:g.kcells])
    z, dz, dzh = c1(g.z), c1(g.dz), c1(g.dzh)
    fn = getattr(lib, "hc_base_state_" + sfx); fn.restype = C.c_int
    bad = fn(*[_p(got[n]) for n in names], _p(thl0), _p(qt0), ct(pbot), C.c_int(g.kstart), C.c_int(g.kend), _p(z), _p(dz), _p(dzh))
    assert bad == 0
    for n in names:
        assert np.array_equal(got[n], ref[n]), (n, got[n], ref[n])
    # the fixed-point form the device kernel uses: the same bits as the serial integration, from scratch, from the previous
    # base state (here: the converged one of a 0.7 K warmer state) and from unusable pressures
    fp = getattr(lib, "hc_base_state_fp_" + sfx); fp.restype = C.c_int
    F = np.zeros(g.kcells, dtype); Fh = np.zeros(g.kcells, dtype); sweeps = C.c_int(0)
    warm_from = O.moist_base_state(g, (thl0 + dtype(0.7)).astype(dtype), qt0, pbot)
    nsw = {}
    for start in ("cold", "warm", "zeros"):
        got2 = {n: (warm_from[n].copy() if start == "warm" else np.zeros(g.kcells, dtype)) for n in names}
        bad = fp(*[_p(got2[n]) for n in names], _p(thl0), _p(qt0), ct(pbot), C.c_int(g.kstart), C.c_int(g.kend), _p(z), _p(dz), _p(dzh),
                 _p(F), _p(Fh), C.c_int(1 if start == "cold" else 0), C.byref(sweeps))
        assert bad == 0
        for n, (lo, hi) in dict(pref=(-1, 1), prefh=(0, 1), rhoref=(0, 0), thvref=(0, 0), exnref=(0, 0), rhorefh=(0, 1), thvrefh=(0, 1), exnrefh=(0, 1)).items():
            sl = slice(g.kstart + lo, g.kend + hi)
            assert np.array_equal(got2[n][sl], ref[n][sl]), (start, n, got2[n][sl], ref[n][sl])
        nsw[start] = sweeps.value
    assert nsw["warm"] <= nsw["cold"] <= 24 and nsw["zeros"] == nsw["cold"], nsw
    ex_fn = getattr(lib, "hc_exner_" + sfx); ex_fn.restype = ct; ex_fn.argtypes = [ct]
    sa = getattr(lib, "hc_sat_adjust_" + sfx); sa.restype = C.c_int
    bu = getattr(lib, "hc_buoyancy_" + sfx)
    nsat = 0
    for k in range(g.kstart, g.kend):
        p = ref["pref"][k]
        ex = dtype(ex_fn(ct(float(p))))
        assert ex == O.moist_exner(dtype, np.array([p], dtype))[0]
        a = np.ascontiguousarray(thl[k].ravel()); b = np.ascontiguousarray(qt[k].ravel())
        o = [np.zeros_like(a) for _ in range(4)]
        assert sa(C.c_long(a.size), _p(a), _p(b), ct(float(p)), ct(float(ex)), *[_p(x) for x in o]) == 0
        r = O.moist_sat_adjust(dtype, a, b, p, ex)
        for x, y, n in zip(o, r, ("ql", "qi", "t", "qs")):
            assert np.array_equal(x, y), (k, n)
        nsat += int((o[0] + o[1] > 0).sum())
        bb = np.zeros_like(a)
        bu(C.c_long(a.size), ct(float(ex)), _p(a), _p(b), _p(o[0]), _p(o[1]), ct(float(ref["thvref"][k])), _p(bb))
        assert np.array_equal(bb, O.moist_buoyancy(dtype, ex, a, b, r[0], r[1], ref["thvref"][k]))
    assert nsat > 100
    a = np.ascontiguousarray(thl[g.kstart].ravel()); b = np.ascontiguousarray(qt[g.kstart].ravel())
    fl1 = np.full_like(a, 8.e-3); fl2 = np.full_like(a, 5.2e-5)
    o1 = np.zeros_like(a); o2 = np.zeros_like(a)
    tv = ref["thvrefh"][g.kstart]
    getattr(lib, "hc_buoyancy_no_ql_" + sfx)(C.c_long(a.size), _p(a), _p(b), ct(float(tv)), _p(o1))
    getattr(lib, "hc_buoyancy_flux_no_ql_" + sfx)(C.c_long(a.size), _p(a), _p(fl1), _p(b), _p(fl2), ct(float(tv)), _p(o2))
    assert np.array_equal(o1, O.moist_buoyancy_no_ql(dtype, a, b, tv))
    assert np.array_equal(o2, O.moist_buoyancy_flux_no_ql(dtype, a, fl1, b, fl2, tv))
