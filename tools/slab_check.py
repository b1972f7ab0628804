#!/usr/bin/env python
"""
Multi-GPU parity check of the y-slab decomposition; run under torchrun, one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/slab_check.py

Every rank owns a slab of ONE global domain.  Checked:
  (1) small grids: one full RK3 step of the slabs == the oracle's single-domain step (rel. L2 within
      BASELINE.json's tolerance), cfl / divergence reductions == the oracle's global values;
  (2) medium grids (warp-FFT and TMA paths): slabs == the same library on a single GPU, bit for bit;
  (3) the 4th-order DNS configuration (advec_4 + diff_4 + pres_4) on slabs == the oracle.
TEST INFRASTRUCTURE: imports oracle/.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from microhh_b200 import dycore as D
from microhh_b200.grid import GridData
from microhh_b200.synthetic import make_case, slab_of
from oracle import oracle as O, step as ostep
from util import TOL, rel_l2, stretched_z


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ok = True

    def slab_run(shape, dtype, anel, stretched, nsteps, ns=1, sizes=(3200., 3200., 3200.)):
        itot, jtot, ktot = shape
        z = stretched_z(ktot, sizes[2]) if stretched else None
        gg = GridData(itot, jtot, ktot, *sizes, 3, 3, 1, dtype, z=z)
        gl = GridData(itot, jtot, ktot, *sizes, 3, 3, 1, dtype, z=z, npy=world, mpicoordy=rank)
        case_g = make_case(gg, seed=2, anelastic=anel, ns=ns)
        case_l = slab_of(case_g, gg, gl)
        ctx = D.Context(gl, lr)
        ctx.set_basestate(case_l["rhoref"], case_l["rhorefh"], case_l["thref"], case_l["threfh"])
        f = D.Fields(ctx, case_l, scalars=case_l["scalars"])
        prm = D.make_params(ns=ns)
        dyc = D.Dycore(ctx, prm)
        for _ in range(nsteps):
            dyc.step(f, 2.0)
        ctx.sync()
        return gg, gl, case_g, ctx, f, prm

    def loc(gl, a_global):
        j0 = gl.jgc + rank*gl.jmax
        return a_global[gl.kstart:gl.kend, j0:j0 + gl.jmax, gl.istart:gl.iend]

    def inter(gl, a):
        return a[gl.kstart:gl.kend, gl.jstart:gl.jend, gl.istart:gl.iend]

    # ---- (1) against the oracle --------------------------------------------------------------
    for dtype in (np.float64, np.float32):
        for shape, anel, st in (((32, 32, 16), False, False), ((48, 24, 12), True, True), ((20, 8*world, 8), True, True)):
            gg, gl, case_g, ctx, f, prm = slab_run(shape, dtype, anel, st, 1)
            g = O.Grid(*shape, 3200., 3200., 3200., 3, 3, 1, dtype, z=stretched_z(shape[2], 3200.) if st else None)
            ostep.dycore_step(g, O.NumpyKernels(g), case_g, ostep.default_params(), 2.0)
            for n in ("u", "v", "w", "th"):
                # the norm is the global one: gather squared sums
                a = inter(gl, f[n].cpu().numpy()).astype(np.float64); b = loc(gl, case_g[n]).astype(np.float64)
                t = torch.tensor([((a - b)**2).sum(), (b**2).sum()], device="cuda", dtype=torch.float64)
                dist.all_reduce(t)
                err = float(torch.sqrt(t[0]/t[1]))
                good = err <= TOL[dtype]
                ok &= good
                if rank == 0:
                    print(f"[oracle] {np.dtype(dtype).name} {shape} P={world} {n}: rel-L2 {err:.2e} {'ok' if good else 'FAIL'}", flush=True)
            # reductions are global
            P2 = O.Pres2(g, case_g["rhoref"], case_g["rhorefh"])
            for n in "uvw":
                O.boundary_cyclic(g, case_g[n])
            div_ref = float(P2.divergence(case_g["u"], case_g["v"], case_g["w"]))
            D.Boundary_cyclic(ctx).exec(f["u"]); D.Boundary_cyclic(ctx).exec(f["v"]); D.Boundary_cyclic(ctx).exec(f["w"])
            div = D.Pres(ctx).check_divergence(f)
            cfl = D.Advec(ctx).get_cfl(f, 2.0)
            cfl_ref = float(O.advec_2i5_cfl(g, case_g["u"], case_g["v"], case_g["w"], 2.0))
            scale = float(np.abs(case_g["u"]).max()/float(g.dx))
            eps = float(np.finfo(dtype).eps)
            good = abs(cfl - cfl_ref) <= 200*TOL[dtype]*cfl_ref and div <= max(10*div_ref, 500*eps*scale)
            ok &= good
            if rank == 0:
                print(f"[oracle] {np.dtype(dtype).name} {shape} cfl {cfl:.6e} (ref {cfl_ref:.6e}) div {div:.2e} (ref {div_ref:.2e}) {'ok' if good else 'FAIL'}", flush=True)
            ctx.close()

    # ---- (2) against the single-GPU path of the same library, bitwise ---------------------------
    for dtype in (np.float64, np.float32):
        for shape in ((128, 64*world, 32), (96, 24*world, 16)):
            gg, gl, case_g, ctx, f, prm = slab_run(shape, dtype, True, True, 2, ns=2)
            case_1 = make_case(gg, seed=2, anelastic=True, ns=2)
            ctx1 = D.Context(gg, lr)
            ctx1.set_basestate(case_1["rhoref"], case_1["rhorefh"], case_1["thref"], case_1["threfh"])
            f1 = D.Fields(ctx1, case_1, scalars=case_1["scalars"])
            dyc1 = D.Dycore(ctx1, D.make_params(ns=2))
            for _ in range(2):
                dyc1.step(f1, 2.0)
            ctx1.sync()
            for n in ("u", "v", "w", "th", "s1", "p"):
                a = inter(gl, f[n].cpu().numpy()); b = loc(gl, f1[n].cpu().numpy())
                t = torch.tensor([float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                good = float(t) == 0.0
                ok &= good
                if rank == 0:
                    print(f"[1gpu]   {np.dtype(dtype).name} {shape} P={world} {n}: max|diff| {float(t):.2e} {'ok' if good else 'FAIL'}", flush=True)
            ctx.close(); ctx1.close()

    # ---- (3) the 4th-order DNS configuration (advec_4 + diff_4 + pres_4, 7-band solve) on slabs, against the oracle ----
    for dtype in (np.float64, np.float32):
        for shape in ((32, 16*world, 16), (48, 12*world, 12)):
            itot, jtot, ktot = shape
            z = stretched_z(ktot, 2.)
            gg = GridData(itot, jtot, ktot, 6., 4., 2., 3, 3, 3, dtype, z=z, order=4)
            gl = GridData(itot, jtot, ktot, 6., 4., 2., 3, 3, 3, dtype, z=z, order=4, npy=world, mpicoordy=rank)
            case_g = make_case(gg, seed=5, noise=0.02)
            ks, ke = gg.kstart, gg.kend
            case_g["w"][:ks+1] = 0; case_g["w"][ke:] = 0
            case_g["th"] = (1. + 0.1*case_g["u"]).astype(dtype)
            for n in ("u", "v"):
                for sfx in ("_bot", "_top", "_gradbot", "_gradtop"):
                    case_g[n + sfx] = np.zeros(gg.shape2d, dtype)
            case_g["th_gradbot"] = np.zeros(gg.shape2d, dtype); case_g["th_gradtop"] = np.zeros(gg.shape2d, dtype)
            case_l = slab_of(case_g, gg, gl)
            ctx = D.Context(gl, lr)
            ones = np.ones(gl.kcells, dtype)
            ctx.set_basestate(ones, ones, 300*ones, 300*ones)
            visc = 1e-3
            f = D.Fields(ctx, case_l, visc=visc, svisc=visc)
            prm = D.make_params(swadvec="4", swdiff="4", swthermo=None, surface_model=False, mbcbot=0, mbctop=0)
            D.Dycore(ctx, prm).step(f, 0.01)
            ctx.sync()
            g = O.Grid(itot, jtot, ktot, 6., 4., 2., 3, 3, 3, dtype, z=z, order=4)
            oprm = ostep.default_params(); oprm.update(swadvec="4", swdiff="4", visc=visc, svisc=visc, mbcbot=0, mbctop=0)
            ostep.dycore_step(g, O.NumpyKernels(g), case_g, oprm, 0.01)
            for n in ("u", "v", "w", "th"):
                a = inter(gl, f[n].cpu().numpy()).astype(np.float64); b = loc(gl, case_g[n]).astype(np.float64)
                t = torch.tensor([((a - b)**2).sum(), (b**2).sum()], device="cuda", dtype=torch.float64)
                dist.all_reduce(t)
                err = float(torch.sqrt(t[0]/t[1]))
                good = err <= 20*TOL[dtype]
                ok &= good
                if rank == 0:
                    print(f"[o4]     {np.dtype(dtype).name} {shape} P={world} {n}: rel-L2 {err:.2e} {'ok' if good else 'FAIL'}", flush=True)
            ctx.close()

    t = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(t)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("SLAB CHECK " + ("PASSED" if int(t) == 0 else "FAILED"), flush=True)
    sys.exit(0 if int(t) == 0 else 1)


if __name__ == "__main__":
    main()
