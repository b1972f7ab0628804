"""
The N > 1 path of Thermo_moist::exec on y slabs, replayed on the CPU with gloo (world_size 2): every rank forms the partial mean
profiles of its slab (sum over its rows / (itot jtot), as moist_mean_profile_kernel does), the ranks all-reduce(sum) them (the
reference's master.sum, src/field3d_operators.cxx:65; the library's ncclAllReduce), every rank integrates the base state and
adds the buoyancy tendency on its slab.  Result == the single-domain oracle: base state to rounding of the split sum, tendencies to
the amplification of that rounding by thvref / |thv - thvref| (see tests/test_zz_gpu_thermo_moist.py).
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
    from util import make_moist_pair
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dtype = np.float64
        g, gd, case, pbot = make_moist_pair(16, 12, 20, dtype)
        thl, qt = case["thl"], case["qt"]
        jmax = g.jtot//world
        rows = slice(g.jstart + rank*jmax, g.jstart + (rank + 1)*jmax)
        n = np.float64(g.itot*g.jtot)
        part = np.empty((2, g.kcells), dtype)
        for i, f in enumerate((thl, qt)):
            part[i] = (f[:, rows, g.istart:g.iend].astype(np.float64).sum(axis=(1, 2))/n).astype(dtype)
        t = torch.from_numpy(part)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        bs = O.moist_base_state(g, part[0], part[1], pbot)
        # buoyancy tendency on my slab only: the oracle kernel on a copy, then cut my rows out
        wt = np.zeros(gd.shape, dtype)
        O.thermo_moist_buoyancy_tend_2nd(g, wt, thl, qt, bs["prefh"], bs["thvrefh"])
        # single-domain reference
        ref_bs = O.moist_base_state(g, O.mean_profile(g, thl), O.mean_profile(g, qt), pbot)
        ref_wt = np.zeros(gd.shape, dtype)
        O.thermo_moist_buoyancy_tend_2nd(g, ref_wt, thl, qt, ref_bs["prefh"], ref_bs["thvrefh"])
        rel = lambda a, b: float(np.sqrt(((a - b)**2).sum()/max((b**2).sum(), 1e-300)))
        sl = slice(g.kstart, g.kend)
        e_bs = max(rel(bs[k][sl], ref_bs[k][sl]) for k in ("pref", "prefh", "thvref", "thvrefh", "rhoref", "rhorefh", "exnref", "exnrefh"))
        e_wt = rel(wt[:, rows], ref_wt[:, rows])
        q.put((rank, e_bs, e_wt, float(np.abs(ref_wt).max())))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_moist_exec_matches_single_domain():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 61
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = []
    while not q.empty():
        res.append(q.get())
    assert len(res) == 2
    for rank, e_bs, e_wt, wmax in res:
        assert e_bs <= 1e-14 and e_wt <= 1e-11 and wmax > 1e-3, (rank, e_bs, e_wt, wmax)
