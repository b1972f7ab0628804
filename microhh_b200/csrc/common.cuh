// mhhb200 -- B200-native dynamical core for MicroHH's RK3 step.
// Shared device-side definitions: grid descriptor, index helpers, finite-difference
// operators (same formulas as reference include/finite_difference.h:33-158, restated).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace mhh {

constexpr int MAX_SCALARS = 8;

// POD copy of the reference's Grid_data scalars (include/grid.h:49-134) plus device
// pointers to the 1-D metric arrays (kcells entries each).
template <typename TF>
struct GridDev
{
    int itot, jtot, ktot;
    int imax, jmax, kmax;
    int igc, jgc, kgc;
    int icells, jcells, kcells;
    int istart, iend, jstart, jend, kstart, kend;
    long long ijcells, ncells;
    TF dx, dy, dxi, dyi;
    TF zsize;
    const TF* z;   const TF* zh;
    const TF* dz;  const TF* dzh;
    const TF* dzi; const TF* dzhi;
    const TF* rhoref;  const TF* rhorefh;
    const TF* thref;   const TF* threfh;
    const TF* dzi4;    const TF* dzhi4;     // 4th-order metrics (NULL on a 2nd-order grid)
};

template <typename TF> __device__ __forceinline__ TF ld(const TF* __restrict__ p) { return __ldg(p); }

// two-element vector type of TF (one 2*sizeof(TF) memory access)
template <typename TF> struct V2T;
template <> struct V2T<double> { typedef double2 type; };
template <> struct V2T<float>  { typedef float2 type; };

// ---- Finite differences ----------------------------------------------------------------
template <typename TF> __device__ __forceinline__ TF interp2(TF a, TF b) { return TF(0.5) * (a + b); }
template <typename TF> __device__ __forceinline__ TF interp4_ws(TF a, TF b, TF c, TF d)
{ return TF(7. / 12.) * (b + c) - TF(1. / 12.) * (a + d); }
template <typename TF> __device__ __forceinline__ TF interp3_ws(TF a, TF b, TF c, TF d)
{ return TF(3. / 12.) * (c - b) - TF(1. / 12.) * (d - a); }
template <typename TF> __device__ __forceinline__ TF interp6_ws(TF a, TF b, TF c, TF d, TF e, TF f)
{ return TF(37. / 60.) * (c + d) - TF(8. / 60.) * (b + e) + TF(1. / 60.) * (a + f); }
template <typename TF> __device__ __forceinline__ TF interp5_ws(TF a, TF b, TF c, TF d, TF e, TF f)
{ return TF(10. / 60.) * (d - c) - TF(5. / 60.) * (e - b) + TF(1. / 60.) * (f - a); }

template <typename TF> __device__ __forceinline__ TF absf(TF a) { return a < TF(0) ? -a : a; }
template <> __device__ __forceinline__ double absf<double>(double a) { return fabs(a); }
template <> __device__ __forceinline__ float absf<float>(float a) { return fabsf(a); }
template <typename TF> __device__ __forceinline__ TF sqrtf_(TF a);
template <> __device__ __forceinline__ double sqrtf_<double>(double a) { return sqrt(a); }
template <> __device__ __forceinline__ float sqrtf_<float>(float a) { return sqrtf(a); }
template <typename TF> __device__ __forceinline__ TF pow2(TF a) { return a * a; }

// Upwind-biased advective face fluxes of the 2i5 scheme:  vel*interp_even - |vel|*interp_odd.
// flux65 in its one-sided form (MHH_UPWIND_F64 / MHH_UPWIND_F32 = 1): vel*i6 - |vel|*i5 is vel*(i6 - i5) for vel >= 0 and
// vel*(i6 + i5) otherwise, and i6 -/+ i5 collapses to the five-point upwind stencil (2, -13, 47, 27, -3)/60 over the cells
// on the upwind side -- 6 floating-point instructions instead of 14 (the six sums / differences of the symmetric form go
// away), at the price of five register selects on the sign of the velocity.  Same real number, rounded differently (well
// inside the 1e-12 / 1e-5 parity budget).  Measured on the B200 at 512^3 (profiles/r02/ab_flux65_*.json): fp32 mom3
// 3.58 -> 3.44 ms per launch (72 instead of 86 registers) -> on; fp64 5.76 -> 7.14 ms (a 64-bit select is two FSEL: 265 fewer
// fp64 instructions but 400 more on the ALU pipe, +12 % issued instructions) -> off.
#ifndef MHH_UPWIND_F64
#define MHH_UPWIND_F64 0
#endif
#ifndef MHH_UPWIND_F32
#define MHH_UPWIND_F32 1
#endif
template <typename TF> struct UpwindForm;
template <> struct UpwindForm<double> { static constexpr bool on = MHH_UPWIND_F64 != 0; };
template <> struct UpwindForm<float> { static constexpr bool on = MHH_UPWIND_F32 != 0; };
__device__ __forceinline__ bool vel_nonneg(double v) { return __double2hiint(v) >= 0; }     // sign bit: an integer compare
__device__ __forceinline__ bool vel_nonneg(float v) { return v >= 0.f; }
template <typename TF> __device__ __forceinline__ TF flux65(TF vel, TF a, TF b, TF c, TF d, TF e, TF f)
{
    if (UpwindForm<TF>::on)
    {
        const bool pos = vel_nonneg(vel);
        const TF q0 = pos ? a : f, q1 = pos ? b : e, q2 = pos ? c : d, q3 = pos ? d : c, q4 = pos ? e : b;
        return vel * (TF(2. / 60.) * q0 + TF(-13. / 60.) * q1 + TF(47. / 60.) * q2 + TF(27. / 60.) * q3 + TF(-3. / 60.) * q4);
    }
    return vel * interp6_ws(a, b, c, d, e, f) - absf(vel) * interp5_ws(a, b, c, d, e, f);
}
template <typename TF> __device__ __forceinline__ TF flux43(TF vel, TF a, TF b, TF c, TF d)
{ return vel * interp4_ws(a, b, c, d) - absf(vel) * interp3_ws(a, b, c, d); }
template <typename TF> __device__ __forceinline__ TF flux2(TF vel, TF a, TF b)
{ return vel * interp2(a, b); }

// Order of the vertical advective flux through level `f` counted on an axis whose first
// interior level is `lo` and whose last is `hi` (faces: lo=kstart, hi=kend; centres for w:
// lo=kstart-1 (virtual), see callers).  0: wall (zero flux), 2, 4 (4th/3rd), 6 (6th/5th).
__device__ __forceinline__ int vorder(int f, int lo, int hi)
{
    const int d = min(f - lo, hi - f);
    return d <= 0 ? 0 : (d == 1 ? 2 : (d == 2 ? 4 : 6));
}

// ---- atomic max for non-negative floating point values -----------------------------------
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v)
{ atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v)); }

template <typename TF>
__device__ __forceinline__ void block_max_to_global(TF v, double* out)
{
    __shared__ double red[32];
    double d = (double)v;
    for (int o = 16; o > 0; o >>= 1) d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nth = blockDim.x * blockDim.y * blockDim.z;
    if ((tid & 31) == 0) red[tid >> 5] = d;
    __syncthreads();
    if (tid < 32)
    {
        d = tid < (nth + 31) / 32 ? red[tid] : 0.;
        for (int o = 16; o > 0; o >>= 1) d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
        if (tid == 0) atomic_max_nonneg(out, d);
    }
}

} // namespace mhh
