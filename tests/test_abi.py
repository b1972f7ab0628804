"""
The C-ABI library loads without a GPU and exports every symbol include/mhhb200.h declares;
the ctypes binding lists exactly those symbols; without a CUDA device a context cannot be
created (no CPU fallback).
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mhhb200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"MHH_API\s+[\w\s\*]+?\b(mhh_\w+)\s*\(", src)))


def test_header_declares_symbols():
    syms = declared_symbols()
    assert "mhh_ctx_create" in syms and "mhh_dycore_substep" in syms and len(syms) >= 25


def test_library_exports_every_declared_symbol():
    from microhh_b200 import capi
    lib = C.CDLL(capi.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_covers_header():
    from microhh_b200 import capi
    assert sorted(capi.SIGNATURES) == declared_symbols()


def test_struct_sizes_match_header_layout():
    """ctypes mirrors of the PODs: spot-check sizes against the C layout rules (LP64)."""
    from microhh_b200 import capi
    assert C.sizeof(capi.GridDesc) == 9*4 + 4 + 3*8 + 6*8 + 4*4 + 2*8    # 9 ints (+pad), 3 doubles, 6 pointers, 4 ints, 2 pointers
    n = capi.MHH_MAX_SCALARS
    assert C.sizeof(capi.FieldsC) == 8*8 + 8 + 2*n*8 + n*8 + 8 + 4*8 + 2*n*8 + 4*8 + 8*8 + 4*n*8 + n*4
    assert C.sizeof(capi.ParamsC) == 5*4 + 4 + 2*8 + 2*4 + 2*n*4


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from microhh_b200 import dycore as D
    from microhh_b200.grid import GridData
    with pytest.raises(D.MhhError):
        D.Context(GridData(16, 16, 8, 1., 1., 1.), 0)
    # and straight through the ABI: create fails with MHH_E_CUDA, no context comes back
    from microhh_b200 import capi
    lib = capi.load()
    gd = GridData(16, 16, 8, 1., 1., 1.)
    keep = [np.ascontiguousarray(getattr(gd, n)) for n in ("z", "zh", "dz", "dzh", "dzi", "dzhi")]
    d = capi.GridDesc(16, 16, 8, 16, 16, 8, 3, 3, 1, 1., 1., 1., *[a.ctypes.data_as(C.c_void_p) for a in keep], 1, 1, 0, 0)
    h = C.c_void_p()
    assert lib.mhh_ctx_create(C.byref(d), 0, 0, C.byref(h)) == -2 and not h


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "microhh_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), fn
                assert "oracle/" not in txt, fn


def test_header_is_plain_c_and_links():
    """include/mhhb200.h is the FFI boundary: it must compile as C99 (no C++ / torch types) and a C program calling the
    GPU-free entry points must link against the library and run (slab layout query; a context cannot be created without a
    CUDA device and must say so)."""
    import subprocess
    import tempfile
    from microhh_b200 import capi
    capi.load()
    src = r"""
#include <stdio.h>
#include "mhhb200.h"
int main(void)
{
    mhh_slab_info s;
    if (mhh_slab_layout(1024, 1024, 1024, 8, 3, &s) != MHH_OK) return 1;
    if (s.nm != 513 || s.jmax != 128 || s.mcl != 64 || s.m_off != 65 + 2*64) return 2;
    if (mhh_slab_layout(1024, 1000, 8, 7, 0, &s) == MHH_OK) return 3;      /* jtot not divisible */
    printf("%d %d %d\n", s.nm, s.mcl, s.m_off);
    return 0;
}
"""
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c"); exe = os.path.join(d, "t")
        with open(c, "w") as fh:
            fh.write(src)
        libdir = os.path.dirname(capi.LIB_PATH)
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{ROOT}/include", c, "-o", exe,
                            f"-L{libdir}", "-lmhhb200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
