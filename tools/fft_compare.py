#!/usr/bin/env python
"""
cuFFT as the PERFORMANCE comparator of the library's Poisson transforms (BASELINE.json north_star).

Times, on the same 512^3 (default) fp64 / fp32 right-hand side:
  * the library's spectral solve (x forward -> y forward + Thomas -> back substitution + y inverse -> x backward, i.e.
    `mhh_pres_fft_roundtrip(solve=1)` minus its test-only staging copies, read from the per-kernel CUDA-event profile);
  * cuFFT through torch: batched D2Z rfft2 + Z2D irfft2 of the same array (two transforms and no solve at all), i.e. the
    lower bound of any cuFFT-based Poisson solver such as the reference's (src/pres.cu:183-285 adds four repack /
    transpose kernels and the tdma kernel on top).
Prints one JSON object; a copy is committed under profiles/r02/.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="512x512x512")
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    from microhh_b200 import dycore as D
    from microhh_b200.grid import GridData
    from microhh_b200.synthetic import make_case
    it, jt, kt = (int(x) for x in args.workload.split("x"))
    dtype = np.float64 if args.dtype == "f64" else np.float32
    gd = GridData(it, jt, kt, 25.*it, 25.*jt, 25.*kt, 3, 3, 1, dtype)
    ones = np.ones(gd.kcells, dtype)
    ctx = D.Context(gd, 0)
    ctx.set_basestate(ones, ones, 300*ones, 300*ones)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    a_in = torch.randn((kt, jt, it), dtype=tdt, device="cuda"); a_out = torch.zeros_like(a_in)
    pres = D.Pres(ctx)
    for _ in range(2):
        pres.fft_roundtrip(a_in, a_out, solve=True)
    ctx.profile_start()
    for _ in range(args.reps):
        pres.fft_roundtrip(a_in, a_out, solve=True)
    prof = ctx.profile_stop()
    ours = {k: v["ms"]/v["n"] for k, v in prof.items()}
    # cuFFT: rfft2 + irfft2 per level batch (the whole array at once: cuFFT batches over k)
    def cufft_once():
        s = torch.fft.rfft2(a_in, dim=(1, 2))
        return torch.fft.irfft2(s, s=(jt, it), dim=(1, 2))
    for _ in range(2):
        cufft_once()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
    tf = tb = 0.
    for _ in range(args.reps):
        e0.record(); s = torch.fft.rfft2(a_in, dim=(1, 2)); e1.record(); o = torch.fft.irfft2(s, s=(jt, it), dim=(1, 2)); e2.record()
        torch.cuda.synchronize()
        tf += e0.elapsed_time(e1); tb += e1.elapsed_time(e2)
    B = np.dtype(dtype).itemsize
    npts = it*jt*kt
    out = {"workload": args.workload, "dtype": args.dtype,
           "ours_ms_per_kernel": ours, "ours_ms_total_solve": sum(ours.values()),
           "cufft_rfft2_ms": tf/args.reps, "cufft_irfft2_ms": tb/args.reps, "cufft_ms_total_no_solve": (tf + tb)/args.reps,
           "pass_ms_at_measured_hbm_peak": npts*B/6469.9e9*1e3,
           "note": "ours includes the tridiagonal solve and writes the ghosted p; cuFFT is transforms only"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
