"""
Restart IO of device fields (Field3d_io<TF>::save_field3d / load_field3d, src/field3d_io.cxx:669-751) through
`mhh_field3d_save` / `mhh_field3d_load`: the file a GPU run writes is byte for byte the file the reference writes (the oracle
restatement is pinned to the compiled reference in tests/test_oracle_vs_ref.py::test_field3d_io_bitexact), either side loads
the other's file, and y slabs write their rows into the one file the way the MPI build's subarray view does.
"""
import numpy as np
import pytest

from util import interior
from oracle import oracle as O
from microhh_b200.grid import GridData

pytestmark = pytest.mark.gpu


def grids(shape, dtype, order=2, npy=1, rank=0):
    gc = (3, 3, 3) if order == 4 else (4, 3, 1)
    g = O.Grid(*shape, 100., 80., 60., *gc, dtype, order=order)
    gd = GridData(*shape, 100., 80., 60., *gc, dtype, order=order, npy=npy, mpicoordy=rank)
    return g, gd


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape,order", [((96, 40, 24), 2), ((64, 1, 16), 2), ((32, 24, 12), 4)])
@pytest.mark.parametrize("offset", [0., 300.])
def test_field3d_save_load(dtype, shape, order, offset, tmp_path):
    import torch
    from microhh_b200 import dycore as D
    g, gd = grids(shape, dtype, order)
    rng = np.random.default_rng(21)
    a = rng.standard_normal(gd.shape).astype(dtype)
    ctx = D.Context(gd, 0)
    io = D.Field3d_io(ctx)
    d_a = torch.from_numpy(a).cuda()
    fg, fo = tmp_path / "u.0000000", tmp_path / "u.oracle"
    assert io.save_field3d(d_a, fg, offset) == 0
    assert O.field3d_save(g, a, str(fo), offset) == 0
    assert fg.read_bytes() == fo.read_bytes()                       # the reference's file, byte for byte
    assert io.save_field3d(d_a, fg, offset) != 0                    # fopen(..., "wbx"): an existing file is an error
    assert np.array_equal(d_a.cpu().numpy(), a)                     # saving does not touch the field
    # load the ORACLE's file into a field full of sevens: interior restored, ghost cells untouched
    d_b = torch.full_like(d_a, 7.)
    assert io.load_field3d(d_b, fo, offset) == 0
    b = d_b.cpu().numpy()
    ref = np.full_like(a, 7.)
    assert O.field3d_load(g, ref, str(fg), offset) == 0             # and the oracle loads the GPU's file
    assert np.array_equal(b, ref)
    if offset == 0.:
        assert np.array_equal(interior(g, b), interior(g, a))       # bitwise identical restarts (src/fields.cxx:1259)
    # errors: missing file, short file
    assert io.load_field3d(d_b, tmp_path / "missing", offset) != 0
    short = tmp_path / "short"; short.write_bytes(fo.read_bytes()[:-8])
    assert io.load_field3d(d_b, short, offset) != 0
    assert b"short" in ctx.lib.mhh_last_error(ctx.h) or b"read" in ctx.lib.mhh_last_error(ctx.h)
    ctx.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("npy", [2, 4])
def test_field3d_slabs_share_one_file(dtype, npy, tmp_path, monkeypatch):
    """Every y slab writes rows [r*jmax, (r+1)*jmax) of each level into the same file (no communicator needed for IO: the
    slab contexts are created on one GPU here); the result is the single-domain file, and each slab reads its part back."""
    import torch
    from microhh_b200 import dycore as D
    monkeypatch.setattr(D.Context, "comm_init", lambda self: None)
    shape = (64, 32, 12)
    g, gd = grids(shape, dtype)
    rng = np.random.default_rng(22)
    a = rng.standard_normal(gd.shape).astype(dtype)
    fo, fs = tmp_path / "th.oracle", tmp_path / "th.0000100"
    assert O.field3d_save(g, a, str(fo)) == 0
    ctxs = []
    for r in reversed(range(npy)):                                  # any order
        _, gl = grids(shape, dtype, npy=npy, rank=r)
        ctx = D.Context(gl, 0)
        sl = np.ascontiguousarray(a[:, r*gl.jmax:r*gl.jmax + gl.jcells, :])
        assert D.Field3d_io(ctx).save_field3d(torch.from_numpy(sl).cuda(), fs) == 0
        ctxs.append((r, gl, ctx, sl))
    assert fs.read_bytes() == fo.read_bytes()
    for r, gl, ctx, sl in ctxs:
        d = torch.full((gl.kcells, gl.jcells, gl.icells), -1., dtype=torch.float64 if dtype == np.float64 else torch.float32, device="cuda")
        assert D.Field3d_io(ctx).load_field3d(d, fo) == 0
        got = d.cpu().numpy()
        assert np.array_equal(got[gl.kstart:gl.kend, gl.jstart:gl.jend, gl.istart:gl.iend], sl[gl.kstart:gl.kend, gl.jstart:gl.jend, gl.istart:gl.iend])
        assert (got[:, :gl.jstart, :] == -1.).all()
        ctx.close()
