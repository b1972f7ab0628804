// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
//
// Shared helpers for the `oracle/_ref/libmhh_ref.so` harness.  Each ref_*.cpp in this
// directory `#include`s ONE translation unit of the reference where it lies under
// /root/reference (nothing is copied into this repository) and exposes the reference's
// anonymous-namespace CPU kernels through a C ABI so that tests and bench.py's
// cpu_baseline leg can call the reference's own arithmetic via ctypes.
#pragma once
#define MHH_EXPORT extern "C" __attribute__((visibility("default")))

// Geometry shared with the stubbed Boundary_cyclic<TF>::exec (ref_stubs.cpp).
struct Ref_geom
{
    int icells, jcells, kcells;
    int istart, iend, jstart, jend, kstart, kend;
    int igc, jgc, jtot;
};
extern Ref_geom ref_geom;

// The reference's own Grid<TF> (ref_grid.cpp): one image per precision, shared by every stand-in object, and the zeroed
// Master image (serial: npx = npy = 1) its `master` reference points to.
void* ref_grid_image(int is_float);
void* ref_master_image();
