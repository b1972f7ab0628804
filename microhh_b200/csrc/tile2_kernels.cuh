// mhhb200 -- TMA / mbarrier primitives and the tile geometry of the TMA-staged, x2-register-blocked z-marching
// tendency kernel (mom3_kernel in tile3_kernels.cuh; fp64 fast path on sm_100a).
//
// One CTA owns an xy tile of 64 x TY columns and marches up a chunk of levels.  Each thread owns
// TWO x-adjacent columns, so every shared-memory access is a 128-bit LDS, the x-face fluxes
// between the two columns are computed once, and the integer/address overhead is halved.
// Horizontal planes (tile + halo) are streamed into a 3-deep shared-memory ring by TMA
// (cp.async.bulk.tensor.3d, one elected thread, completion on an mbarrier): no per-thread staging
// instructions or registers, out-of-bounds halo columns are zero-filled by the hardware.
// The own columns live in sliding register windows; vertical fluxes are computed once per face and
// carried to the next level.  Divisions by the base-state density are replaced by per-level
// reciprocals held in shared memory (B200's fp64 pipe, not HBM, bounds these kernels).
//
// Arithmetic: flux form of reference src/advec_2i5.cxx:151-728, include/diff_kernels.h:144-484,
// src/thermo_dry.cxx:165-179 (same formulas as stencil_kernels.cuh; sums are re-associated at the
// 1-ulp level where two columns share a term).
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "stencil_kernels.cuh"
#include "tile_kernels.cuh"

namespace mhh {

// ---- TMA / mbarrier primitives ---------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory"); }

__device__ __forceinline__ void mbar_fence_init()
{ asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory"); }

__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    do
    {
        // the suspend-time hint lets a waiting warp sleep instead of spinning in the issue slots of the working warps
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
    } while (!ok);
}

// 3-D tiled TMA load: box (bx, by, 1) at element coordinates (x, y, z); out-of-range elements are zero-filled.
__device__ __forceinline__ void tma_load_3d(unsigned smem_dst, const CUtensorMap* tmap, unsigned bar, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                 :: "r"(smem_dst), "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
}

// L2 prefetch of a box (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* tmap, int x, int y, int z)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];\n"
                 :: "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(x), "r"(y), "r"(z) : "memory");
}


// ---- tile geometry ---------------------------------------------------------------------------
constexpr int T2_W  = 64;                 // tile width  (2 columns per lane)
// x halo HL (3 or 4).  The TMA box origin istart + 64 bx - HL must be 16-byte aligned in global memory (probed on B200: an odd
// fp64 x coordinate raises "illegal instruction", a negative even one is fine and zero-filled).
//   HL = 4 (igc = 4: what the adapters ask Grid for, set_minimum_ghost_cells; fp64 and fp32): the thread's own pair sits at
//     an EVEN column of the staged plane and at a 16-byte (fp32: 8-byte) aligned global address: one vector LDS / LDG / STG
//     per pair.
//   HL = 3 (fp64 fields with the reference's minimum igc = 3): the interior starts at an odd element, the own pair sits at
//     an ODD shared-memory column and is read as two 8-byte loads; rows along x are read as aligned 16-byte pairs starting
//     one column early.  24 % more shared-memory wavefronts than HL = 4 (ncu), and 8-byte global accesses.
constexpr int t2_px(int hl) { return T2_W + 2 * hl; }      // 70 / 72: plane pitch = TMA box width (multiples of 16 bytes)
constexpr int T2_H  = 3;                  // y halo
// plane size in elements, padded so that every plane starts on a 128-byte boundary (TMA destination alignment)
constexpr int t2_plane(int ty, int elem, int hl) { return (t2_px(hl) * (ty + 2 * T2_H) * elem + 127) / 128 * 128 / elem; }
constexpr int t2_box_bytes(int ty, int elem, int hl) { return t2_px(hl) * (ty + 2 * T2_H) * elem; }
// the halo a grid can use: 4 when istart - 4 is 16-byte aligned, else 3 for fp64 when istart - 3 is; 0 = no TMA path
inline int t2_pick_hl(int igc, int elem)
{
    if (igc >= 4 && ((igc - 4) * elem) % 16 == 0) return 4;
    if (elem == 8 && igc >= 3 && ((igc - 3) * elem) % 16 == 0) return 3;
    return 0;
}

} // namespace mhh
