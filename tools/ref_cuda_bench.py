#!/usr/bin/env python
"""
The reference's single-GPU CUDA kernels timed on the B200 "for context" (BASELINE.json north_star).

oracle/_ref/libmhh_refcuda.so holds MicroHH's own kernels of the hot path -- advec_2i5 (u, v, w, s), calc_strain2 + evisc,
diff_uvw + diff_c, pres_2 (pres_in, solve_in, tdma, solve_out, pres_out), rk3 -- compiled for sm_100a from the headers under
/root/reference/include and launched through the reference's own launcher with its default block sizes.  They run here on
the same synthetic 512^3 fields bench.py uses; the reference's FFT stage is cuFFT, timed through torch.fft (rfft2 + irfft2;
the reference adds four repack / transpose kernels around it, src/pres.cu:287-495, which are not included: optimistic for
the reference).  Prints one JSON object: per-kernel milliseconds, the reference's sub-step as the sum over its launch list
(src/model.cxx:356-504 for drycblles-type S = 1) and this library's sub-step measured in the same process.
This is measurement infrastructure: nothing here is imported by the product.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

PTRS = ["u", "v", "w", "s", "ut", "vt", "wt", "st", "evisc", "p", "tmp1", "tmp2", "n2",
        "fluxbotu", "fluxtopu", "fluxbotv", "fluxtopv", "fluxbots", "fluxtops", "dudz", "dvdz", "dbdz", "z0m",
        "z", "dz", "dzi", "dzhi", "rhoref", "rhorefh", "rhorefi", "rhorefhi", "mlen", "a", "c", "bmati", "bmatj"]


class RefArgs(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("istart", "iend", "jstart", "jend", "kstart", "kend", "icells", "ijcells",
                                         "imax", "jmax", "kmax", "igc", "jgc", "kgc")]
                + [(n, C.c_double) for n in ("dxi", "dyi", "dt", "tPri", "visc")]
                + [(n, C.c_void_p) for n in PTRS])


KERNELS = ["advec_u", "advec_v", "advec_w", "advec_s", "calc_strain2", "evisc", "diff_uvw", "diff_c",
           "pres_in", "solve_in", "tdma", "solve_out", "pres_out", "rk3"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="512x512x512")
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    from microhh_b200 import dycore as D
    from microhh_b200.grid import GridData
    from microhh_b200.synthetic import fill_fields_device
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmhh_refcuda.so"))
    assert lib.refcuda_args_size() == C.sizeof(RefArgs), (lib.refcuda_args_size(), C.sizeof(RefArgs))
    it, jt, kt = (int(x) for x in args.workload.split("x"))
    dtype = np.float64 if args.dtype == "f64" else np.float32
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    gd = GridData(it, jt, kt, 25.*it, 25.*jt, 25.*kt, 3, 3, 1, dtype)
    ctx = D.Context(gd, 0)
    f = D.Fields(ctx, None)
    prof = fill_fields_device(f, gd, noise=0.01)
    ctx.set_basestate(prof["rhoref"], prof["rhorefh"], prof["thref"], prof["threfh"])
    dev = f["u"].device
    t1 = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(dtype))).to(dev)
    keep = {}
    keep["tmp1"] = torch.zeros_like(f["u"]); keep["tmp2"] = torch.zeros_like(f["u"]); keep["n2"] = torch.full_like(f["u"], 1e-5)
    for n in ("z", "dz", "dzi", "dzhi"):
        keep[n] = t1(getattr(gd, n))
    keep["rhoref"] = t1(prof["rhoref"]); keep["rhorefh"] = t1(prof["rhorefh"])
    keep["rhorefi"] = t1(1./prof["rhoref"]); keep["rhorefhi"] = t1(1./prof["rhorefh"])
    keep["mlen"] = t1((0.23*np.cbrt(float(gd.dx)*float(gd.dy)*gd.dz.astype(np.float64)))**2)
    keep["a"] = t1(np.ones(kt)); keep["c"] = t1(np.ones(kt))
    keep["bmati"] = t1(-np.abs(np.linspace(0, 1, it))); keep["bmatj"] = t1(-np.abs(np.linspace(0, 1, jt)) - 1e-3)
    r = RefArgs()
    r.istart, r.iend, r.jstart, r.jend, r.kstart, r.kend = gd.istart, gd.iend, gd.jstart, gd.jend, gd.kstart, gd.kend
    r.icells, r.ijcells, r.imax, r.jmax, r.kmax = gd.icells, gd.icells*gd.jcells, gd.imax, gd.jmax, gd.kmax
    r.igc, r.jgc, r.kgc = gd.igc, gd.jgc, gd.kgc
    r.dxi, r.dyi, r.dt, r.tPri, r.visc = 1./float(gd.dx), 1./float(gd.dy), 1.0, 3.0, 1e-5
    src = {"u": f["u"], "v": f["v"], "w": f["w"], "s": f["th"], "ut": f["ut"], "vt": f["vt"], "wt": f["wt"], "st": f["tht"],
           "evisc": f["evisc"], "p": f["p"], "fluxbotu": f["u_fluxbot"], "fluxtopu": f["u_fluxtop"], "fluxbotv": f["v_fluxbot"],
           "fluxtopv": f["v_fluxtop"], "fluxbots": f["th_fluxbot"], "fluxtops": f["th_fluxtop"], "dudz": f["dudz_mo"],
           "dvdz": f["dvdz_mo"], "dbdz": f["dbdz_mo"], "z0m": f["z0m"]}
    src.update(keep)
    for n in PTRS:
        setattr(r, n, src[n].data_ptr())
    # valid halos / viscosity so that the kernels see sane numbers
    dyc = D.Dycore(ctx, D.make_params())
    dyc.substep_pre(f)
    ctx.sync()
    torch.cuda.synchronize()
    ms = {}
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    s_legacy = torch.cuda.default_stream()
    with torch.cuda.stream(s_legacy):
        for w, name in enumerate(KERNELS):
            for _ in range(2):
                rc = lib.refcuda_launch(w, int(dtype == np.float64), C.byref(r))
                assert rc == 0, (name, rc)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.reps):
                lib.refcuda_launch(w, int(dtype == np.float64), C.byref(r))
            e1.record(); torch.cuda.synchronize()
            ms[name] = e0.elapsed_time(e1)/args.reps
            for n in ("ut", "vt", "wt", "tht"):
                f[n].zero_()
        # cuFFT stage of the reference's pres_2 (transforms only)
        a_in = torch.randn((kt, jt, it), dtype=tdt, device=dev)
        for _ in range(2):
            torch.fft.irfft2(torch.fft.rfft2(a_in, dim=(1, 2)), s=(jt, it), dim=(1, 2))
        torch.cuda.synchronize(); e0.record()
        for _ in range(args.reps):
            torch.fft.irfft2(torch.fft.rfft2(a_in, dim=(1, 2)), s=(jt, it), dim=(1, 2))
        e1.record(); torch.cuda.synchronize()
        ms["cufft_rfft2_irfft2"] = e0.elapsed_time(e1)/args.reps
        del a_in
    # the reference's launch list of one sub-step, S = 1 (src/model.cxx:356-504): rk3 for u, v, w, th
    ref_sub = (ms["advec_u"] + ms["advec_v"] + ms["advec_w"] + ms["advec_s"] + ms["calc_strain2"] + ms["evisc"]
               + ms["diff_uvw"] + ms["diff_c"] + ms["pres_in"] + ms["cufft_rfft2_irfft2"] + ms["solve_in"] + ms["tdma"]
               + ms["solve_out"] + ms["pres_out"] + 4*ms["rk3"])
    # this library, same process, same fields
    fill_fields_device(f, gd, noise=0.01)
    for n in ("ut", "vt", "wt", "tht"):
        f[n].zero_()
    for _ in range(2):
        dyc.step(f, 1.0)
    ctx.sync(); torch.cuda.synchronize()
    ctx.profile_start()
    e0.record()
    for _ in range(args.reps):
        dyc.step(f, 1.0)
    e1.record(); torch.cuda.synchronize()
    ours = ctx.profile_stop()
    ours_sub = e0.elapsed_time(e1)/args.reps/3
    out = {"workload": args.workload, "dtype": args.dtype, "reference_cuda_kernels_ms": ms,
           "reference_substep_ms_sum_of_kernels": ref_sub,
           "not_included_for_the_reference": "boundary_cyclic halos (>= 8 launches), the four cuFFT repack/transpose kernels, the N2 kernel, set_ghost_cells, cudaMemcpy of p",
           "ours_substep_ms": ours_sub, "ours_kernels_ms_per_substep": {k: v["ms"]/args.reps/3 for k, v in sorted(ours.items(), key=lambda kv: -kv[1]["ms"])},
           "ratio_reference_over_ours": ref_sub/ours_sub}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
