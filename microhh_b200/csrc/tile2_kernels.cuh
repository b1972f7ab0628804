// mhhb200 -- TMA-staged, x2-register-blocked z-marching tile kernels (fp64 fast path on sm_100a).
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

template <typename TF> struct V2T;
template <> struct V2T<double> { typedef double2 type; };
template <> struct V2T<float>  { typedef float2 type; };

// ---- tile geometry ---------------------------------------------------------------------------
constexpr int T2_W  = 64;                 // tile width  (2 columns per lane)
constexpr int T2_HL = 3;                  // x halo.  TMA needs a 16-byte aligned box origin: istart - 3 = 0 is even, so the
                                          // thread's own pair sits at an ODD shared-memory column and is read as two 8-byte
                                          // loads; rows along x are read as aligned 16-byte pairs starting one column early.
constexpr int T2_PX = T2_W + 2 * T2_HL;   // 70: plane pitch (= TMA box width; 560 bytes, a multiple of 16)
constexpr int T2_H  = 3;                  // y halo
constexpr int T2_RING = 3;
// plane size in elements, padded so that every plane starts on a 128-byte boundary (TMA destination alignment)
constexpr int t2_plane(int ty, int elem = 8) { return (T2_PX * (ty + 2 * T2_H) * elem + 127) / 128 * 128 / elem; }
constexpr int t2_box_bytes(int ty, int elem = 8) { return T2_PX * (ty + 2 * T2_H) * elem; }

template <typename TF>
struct Mom2Args
{
    MomArgs<TF> m;
    int kchunk;
    int prefetch;      // L2 prefetch distance in levels (0 = off)
};

// shared-memory layout: [0,128) mbarriers | planes [field][ring] | level profiles
inline size_t mom2_smem(size_t elem, int kchunk, int ty)
{ return 128 + ((size_t)4 * T2_RING * t2_plane(ty, (int)elem) + (size_t)8 * (kchunk + 3)) * elem + 128; }

template <typename TF, bool SURFACE, bool BUOY, int TY>
__global__ void __launch_bounds__(32 * TY, (TY <= 8) ? 2 : 1)
mom2_kernel(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_v,
            const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_e,
            const __grid_constant__ CUtensorMap tm_ut, const __grid_constant__ CUtensorMap tm_vt,
            const __grid_constant__ CUtensorMap tm_wt, const __grid_constant__ CUtensorMap tm_th,
            const Mom2Args<TF> args, const GridDev<TF> g)
{
    typedef typename V2T<TF>::type V2;
    extern __shared__ unsigned char smem_raw[];
    // 128-byte aligned base (TMA destination alignment)
    unsigned char* sbase = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sbase);
    TF* sm = reinterpret_cast<TF*>(sbase + 128);
    constexpr int PLANE = t2_plane(TY, (int)sizeof(TF)), P = T2_PX, NT = 32 * TY;
    constexpr unsigned PLANE_BYTES = PLANE * sizeof(TF), BOX_BYTES = t2_box_bytes(TY, (int)sizeof(TF));

    const MomArgs<TF>& a = args.m;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = g.istart + blockIdx.x * T2_W + 2 * tx;            // first of the two columns
    const int j = g.jstart + blockIdx.y * TY + ty;
    const int gi0 = g.istart + blockIdx.x * T2_W - T2_HL;           // even (igc = 3): 16-byte aligned box origin
    const int gj0 = g.jstart + blockIdx.y * TY - T2_H;
    const bool active = (i + 1 < g.iend) && (j < g.jend);
    const int ks = g.kstart, ke = g.kend;
    const int kc0 = ks + blockIdx.z * args.kchunk;
    const int kc1 = min(ke, kc0 + args.kchunk);
    const long long jj = g.icells, kk = g.ijcells;
    const int ic = min(i, g.iend - 2), jc = min(j, g.jend - 1);     // clamp inactive threads in bounds
    const long long ij = ic + jc * jj;
    const int sidx = (ty + T2_H) * P + T2_HL + 2 * tx;              // odd column: sidx - 1 is 16-byte aligned
    const TF dxi = g.dxi, dyi = g.dyi, visc = a.visc;
    const TF q = TF(0.25);

    // ---- per-level profiles of this chunk (levels kc0-1 .. kc1+1) ----
    const int k0 = kc0 - 1;
    TF* prof = sm + 4 * T2_RING * PLANE;
    const int nlev = args.kchunk + 3;
    TF* p_rho = prof; TF* p_rhoh = prof + nlev; TF* p_rdzi = prof + 2 * nlev; TF* p_rdzhi = prof + 3 * nlev;
    TF* p_dzhi = prof + 4 * nlev; TF* p_gth = prof + 5 * nlev; TF* p_thh = prof + 6 * nlev; TF* p_dzi = prof + 7 * nlev;
    for (int t = threadIdx.x; t < nlev; t += NT)
    {
        const int lev = min(max(k0 + t, 0), g.kcells - 1);
        const TF rho = g.rhoref[lev], rhoh = g.rhorefh[lev];
        p_rho[t] = rho; p_rhoh[t] = rhoh;
        p_rdzi[t] = g.dzi[lev] / rho;            // (.)/rhoref[k]*dzi[k]   -> (.)*rdzi
        p_rdzhi[t] = g.dzhi[lev] / rhoh;         // (.)/rhorefh[k]*dzhi[k] -> (.)*rdzhi
        p_dzhi[t] = g.dzhi[lev];
        p_gth[t] = BUOY ? TF(GRAV) / g.threfh[lev] : TF(0);
        p_thh[t] = BUOY ? g.threfh[lev] : TF(0);
        p_dzi[t] = g.dzi[lev];
    }

    // ---- TMA pipeline ----
    const unsigned bar0 = smem_u32(bars);
    const unsigned pl0 = smem_u32(sm);
    if (threadIdx.x == 0)
    {
        for (int s = 0; s < T2_RING; ++s) mbar_init(bar0 + 8 * s, 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int slot, int lev) {
        const unsigned bar = bar0 + 8 * slot;
        mbar_expect_tx(bar, 4 * BOX_BYTES);
        tma_load_3d(pl0 + (0 * T2_RING + slot) * PLANE_BYTES, &tm_u, bar, gi0, gj0, lev);
        tma_load_3d(pl0 + (1 * T2_RING + slot) * PLANE_BYTES, &tm_v, bar, gi0, gj0, lev);
        tma_load_3d(pl0 + (2 * T2_RING + slot) * PLANE_BYTES, &tm_w, bar, gi0, gj0, lev);
        tma_load_3d(pl0 + (3 * T2_RING + slot) * PLANE_BYTES, &tm_e, bar, gi0, gj0, lev);
    };
    if (threadIdx.x == 0) { issue(0, k0); issue(1, k0 + 1); }
    unsigned phase = 0;       // bit s = parity to wait for on slot s
    int s0 = 0;

    auto colload = [&](const TF* __restrict__ fld, int lev, int c) -> TF {
        return (lev >= 0 && lev < g.kcells) ? fld[ij + c + (long long)lev * kk] : TF(0);
    };
    // register windows (per column c): ua = u[k-2], ub = u[k-1], uc = u[k+2], ud = u[k+3]  (u[k], u[k+1] come from the planes)
    //                                  wa = w[k-1], wb = w[k],   wc = w[k+3], wd = w[k+4]  (w[k+1], w[k+2]: plane k+1 / register we)
    TF ua[2], ub[2], uc[2], ud[2], va[2], vb[2], vc[2], vd[2], wa[2], wb[2], we[2], wc[2], wd[2];
#pragma unroll
    for (int c = 0; c < 2; ++c)
    {
        ua[c] = colload(a.u, k0 - 2, c); ub[c] = colload(a.u, k0 - 1, c); uc[c] = colload(a.u, k0 + 2, c); ud[c] = colload(a.u, k0 + 3, c);
        va[c] = colload(a.v, k0 - 2, c); vb[c] = colload(a.v, k0 - 1, c); vc[c] = colload(a.v, k0 + 2, c); vd[c] = colload(a.v, k0 + 3, c);
        wa[c] = colload(a.w, k0 - 1, c); wb[c] = colload(a.w, k0, c); we[c] = colload(a.w, k0 + 2, c);
        wc[c] = colload(a.w, k0 + 3, c); wd[c] = colload(a.w, k0 + 4, c);
    }
    TF thk[2] = {TF(0), TF(0)};
    if (BUOY) { thk[0] = colload(a.th, k0, 0); thk[1] = colload(a.th, k0, 1); }

    // carried vertical flux (diffusive - advective) through the bottom face (u, v) / bottom centre (w)
    TF gu[2] = {0, 0}, gv[2] = {0, 0}, gw[2] = {0, 0};

    mbar_wait(bar0 + 8 * 0, 0); phase ^= 1u;         // plane k0

    for (int k = k0; k < kc1; ++k)
    {
        const int s1 = (s0 == T2_RING - 1) ? 0 : s0 + 1;
        const int s2 = (s1 == T2_RING - 1) ? 0 : s1 + 1;
        __syncthreads();                              // everyone is done with plane k-1 (slot s2)
        if (threadIdx.x == 0)
        {
            if (k + 2 <= kc1) issue(s2, k + 2);   // only planes that will be consumed
            // pull what the per-thread loads of the NEXT iterations touch into L2: the leading window levels of
            // u, v, w (also the first DRAM touch of the planes staged three iterations later), the tendencies, th
            if (args.prefetch)
            {
                const int pu = k + 1 + args.prefetch + 3;
                if (pu < g.kcells) { tma_prefetch_3d(&tm_u, gi0, gj0, pu); tma_prefetch_3d(&tm_v, gi0, gj0, pu); }
                if (pu + 1 < g.kcells) tma_prefetch_3d(&tm_w, gi0, gj0, pu + 1);
                const int pt = k + args.prefetch;
                if (pt + 1 <= kc1)
                {
                    tma_prefetch_3d(&tm_ut, gi0 + 2, gj0 + T2_H, pt); tma_prefetch_3d(&tm_vt, gi0 + 2, gj0 + T2_H, pt);
                    tma_prefetch_3d(&tm_wt, gi0 + 2, gj0 + T2_H, pt + 1);
                    if (BUOY) tma_prefetch_3d(&tm_th, gi0 + 2, gj0 + T2_H, pt + 1);
                }
            }
        }

        const bool store = (k >= kc0);
        const int f = k + 1;
        const bool st_uv = store && active;
        const bool st_w = st_uv && f < ke;
        const long long o_k = ij + (long long)k * kk, o_f = o_k + kk;

        mbar_wait(bar0 + 8 * s1, (phase >> s1) & 1u); phase ^= (1u << s1);     // plane k+1 has landed

        const TF* __restrict__ U0 = sm + (0 * T2_RING + s0) * PLANE + sidx;
        const TF* __restrict__ U1 = sm + (0 * T2_RING + s1) * PLANE + sidx;
        const TF* __restrict__ V0 = sm + (1 * T2_RING + s0) * PLANE + sidx;
        const TF* __restrict__ V1 = sm + (1 * T2_RING + s1) * PLANE + sidx;
        const TF* __restrict__ W1 = sm + (2 * T2_RING + s1) * PLANE + sidx;
        const TF* __restrict__ E0 = sm + (3 * T2_RING + s0) * PLANE + sidx;
        const TF* __restrict__ E1 = sm + (3 * T2_RING + s1) * PLANE + sidx;
        auto LD2 = [](const TF* p) -> V2 { return *reinterpret_cast<const V2*>(p); };
        // Row loaders.  The own pair starts at an odd column, so rows are read as ALIGNED 16-byte pairs starting one
        // column early: x[n] below is the value at column offset n from the first own column.
        //   row10: x[-5..6] -> array index n+5 (x[-4..5] are used)      6 vector loads
        //   row6 : x[-3..4] -> array index n+3 (x[-2..3] are used)      4 vector loads
        //   row4 : x[-1..4] -> array index n+1 (x[0..3]  are used)      3 vector loads
        auto row10 = [&](const TF* p, TF (&x)[12]) {
#pragma unroll
            for (int n = 0; n < 6; ++n) { const V2 t = LD2(p + 2 * n - 5); x[2 * n] = t.x; x[2 * n + 1] = t.y; } };
        auto row6 = [&](const TF* p, TF (&x)[8]) {
#pragma unroll
            for (int n = 0; n < 4; ++n) { const V2 t = LD2(p + 2 * n - 3); x[2 * n] = t.x; x[2 * n + 1] = t.y; } };
        auto row4 = [&](const TF* p, TF (&x)[6]) {
#pragma unroll
            for (int n = 0; n < 3; ++n) { const V2 t = LD2(p + 2 * n - 1); x[2 * n] = t.x; x[2 * n + 1] = t.y; } };
        auto col7 = [&](const TF* p, TF (&y)[7][2]) {
#pragma unroll
            for (int d = -3; d <= 3; ++d) { y[d + 3][0] = p[d * P]; y[d + 3][1] = p[d * P + 1]; } };
        auto pair = [&](const TF* p, TF (&x)[2]) { x[0] = p[0]; x[1] = p[1]; };
        // pair sums along x of a viscosity row: S[m] = E[m-1] + E[m], m = 0..2, from x[-2..3]
        auto psum = [&](const TF (&e)[8], TF (&sum)[3]) {
#pragma unroll
            for (int m = 0; m < 3; ++m) sum[m] = e[m + 2] + e[m + 3]; };

        const int pl = k - k0;
        const TF rhoh_f = p_rhoh[pl + 1];
        const TF rdzi_k = p_rdzi[pl], dzhi_f = p_dzhi[pl + 1];
        const int of = vorder(f, ks, ke);
        // the compiler must not hoist the loads of a later component above an earlier one (register pressure)
#define BLOCK_FENCE() asm volatile("" ::: "memory")

#define X10(a, n) a[(n) + 5]
#define X6(a, n) a[(n) + 3]
#define X4(a, n) a[(n) + 1]
        // =========================================================== u at cell k
        {
            const TF un0 = colload(a.u, k + 4, 0), un1 = colload(a.u, k + 4, 1);
            TF old0 = 0, old1 = 0;
            if (st_uv) { old0 = a.ut[o_k]; old1 = a.ut[o_k + 1]; }
            TF ux[12], w6[8], e0[8], e1[8], s0r[3], s1r[3], U1R[2];
            row10(U0, ux); row6(W1, w6); row6(E0, e0); row6(E1, e1); pair(U1, U1R);
            psum(e0, s0r); psum(e1, s1r);
            // x-face fluxes between u(m-1) and u(m), m = 0..2 (advective and diffusive)
            TF fx[3], dx_[3];
#pragma unroll
            for (int m = 0; m < 3; ++m)
            {
                fx[m] = flux65(interp2(X10(ux, m - 1), X10(ux, m)), X10(ux, m - 3), X10(ux, m - 2), X10(ux, m - 1), X10(ux, m), X10(ux, m + 1), X10(ux, m + 2));
                dx_[m] = (X6(e0, m - 1) + visc) * (X10(ux, m) - X10(ux, m - 1)) * dxi;
            }
            TF gt[2];
#pragma unroll
            for (int c = 0; c < 2; ++c)
            {
                const TF uk = X10(ux, c), uk1 = U1R[c];
                const TF ft_a = rhoh_f * vflux_col<TF>(of, interp2(X6(w6, c - 1), X6(w6, c)), ua[c], ub[c], uk, uk1, uc[c], ud[c]);
                TF ft_d;
                if (SURFACE && f == ks) ft_d = -rhoh_f * a.u_fluxbot[ij + c];
                else if (SURFACE && f == ke) ft_d = -rhoh_f * a.u_fluxtop[ij + c];
                else
                {
                    const TF evisct = q * (s0r[c] + s1r[c]) + visc;
                    ft_d = rhoh_f * evisct * ((uk1 - uk) * dzhi_f + (X6(w6, c) - X6(w6, c - 1)) * dxi);
                }
                gt[c] = ft_d - ft_a;
            }
            if (st_uv)
            {
                TF uy[7][2], v6[8], vp6[8], em[8], ep[8], s0m[3], s0p[3];
                col7(U0, uy); row6(V0, v6); row6(V0 + P, vp6); row6(E0 - P, em); row6(E0 + P, ep);
                psum(em, s0m); psum(ep, s0p);
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    const TF fn = flux65(interp2(X6(vp6, c - 1), X6(vp6, c)), uy[1][c], uy[2][c], uy[3][c], uy[4][c], uy[5][c], uy[6][c]);
                    const TF fs = flux65(interp2(X6(v6, c - 1), X6(v6, c)), uy[0][c], uy[1][c], uy[2][c], uy[3][c], uy[4][c], uy[5][c]);
                    const TF eviscn = q * (s0r[c] + s0p[c]) + visc;
                    const TF eviscs = q * (s0m[c] + s0r[c]) + visc;
                    const TF d = (dx_[c + 1] - dx_[c]) * TF(2.) * dxi
                               + (eviscn * ((uy[4][c] - uy[3][c]) * dyi + (X6(vp6, c) - X6(vp6, c - 1)) * dxi)
                                - eviscs * ((uy[3][c] - uy[2][c]) * dyi + (X6(v6, c) - X6(v6, c - 1)) * dxi)) * dyi;
                    const TF tu = -(fx[c + 1] - fx[c]) * dxi - (fn - fs) * dyi + d + (gt[c] - gu[c]) * rdzi_k;
                    a.ut[o_k + c] = (c == 0 ? old0 : old1) + tu;
                }
            }
            gu[0] = gt[0]; gu[1] = gt[1];
            ua[0] = ub[0]; ua[1] = ub[1]; ub[0] = X10(ux, 0); ub[1] = X10(ux, 1);
            uc[0] = ud[0]; uc[1] = ud[1]; ud[0] = un0; ud[1] = un1;
        }
        BLOCK_FENCE();
        // =========================================================== v at cell k
        {
            const TF vn0 = colload(a.v, k + 4, 0), vn1 = colload(a.v, k + 4, 1);
            TF old0 = 0, old1 = 0;
            if (st_uv) { old0 = a.vt[o_k]; old1 = a.vt[o_k + 1]; }
            TF vx[12], e0[8], em[8], V1R[2], W1R[2], W1M[2], E1R[2], E1M[2];
            row10(V0, vx); row6(E0, e0); row6(E0 - P, em);
            pair(V1, V1R); pair(W1, W1R); pair(W1 - P, W1M); pair(E1, E1R); pair(E1 - P, E1M);
            TF gt[2];
#pragma unroll
            for (int c = 0; c < 2; ++c)
            {
                const TF vk = X10(vx, c), vk1 = V1R[c];
                const TF ft_a = rhoh_f * vflux_col<TF>(of, interp2(W1M[c], W1R[c]), va[c], vb[c], vk, vk1, vc[c], vd[c]);
                TF ft_d;
                if (SURFACE && f == ks) ft_d = -rhoh_f * a.v_fluxbot[ij + c];
                else if (SURFACE && f == ke) ft_d = -rhoh_f * a.v_fluxtop[ij + c];
                else
                {
                    const TF evisct = q * ((X6(em, c) + X6(e0, c)) + (E1M[c] + E1R[c])) + visc;
                    ft_d = rhoh_f * evisct * ((vk1 - vk) * dzhi_f + (W1R[c] - W1M[c]) * dyi);
                }
                gt[c] = ft_d - ft_a;
            }
            if (st_uv)
            {
                TF u4[6], um4[6], vy[7][2], s0r[3], s0m[3];
                row4(U0, u4); row4(U0 - P, um4); col7(V0, vy);
                psum(e0, s0r); psum(em, s0m);
                TF fx[3], dx_[3];
#pragma unroll
                for (int m = 0; m < 3; ++m)
                {
                    fx[m] = flux65(interp2(X4(um4, m), X4(u4, m)), X10(vx, m - 3), X10(vx, m - 2), X10(vx, m - 1), X10(vx, m), X10(vx, m + 1), X10(vx, m + 2));
                    const TF eviscc = q * (s0m[m] + s0r[m]) + visc;
                    dx_[m] = eviscc * ((X10(vx, m) - X10(vx, m - 1)) * dxi + (X4(u4, m) - X4(um4, m)) * dyi);
                }
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    const TF fn = flux65(interp2(vy[3][c], vy[4][c]), vy[1][c], vy[2][c], vy[3][c], vy[4][c], vy[5][c], vy[6][c]);
                    const TF fs = flux65(interp2(vy[2][c], vy[3][c]), vy[0][c], vy[1][c], vy[2][c], vy[3][c], vy[4][c], vy[5][c]);
                    const TF d = (dx_[c + 1] - dx_[c]) * dxi
                               + ((X6(e0, c) + visc) * (vy[4][c] - vy[3][c]) * dyi - (X6(em, c) + visc) * (vy[3][c] - vy[2][c]) * dyi) * TF(2.) * dyi;
                    const TF tv = -(fx[c + 1] - fx[c]) * dxi - (fn - fs) * dyi + d + (gt[c] - gv[c]) * rdzi_k;
                    a.vt[o_k + c] = (c == 0 ? old0 : old1) + tv;
                }
            }
            gv[0] = gt[0]; gv[1] = gt[1];
            va[0] = vb[0]; va[1] = vb[1]; vb[0] = X10(vx, 0); vb[1] = X10(vx, 1);
            vc[0] = vd[0]; vc[1] = vd[1]; vd[0] = vn0; vd[1] = vn1;
        }
        BLOCK_FENCE();
        // =========================================================== w at face f = k+1
        {
            const TF wn0 = colload(a.w, k + 5, 0), wn1 = colload(a.w, k + 5, 1);
            TF th1[2] = {TF(0), TF(0)};
            if (BUOY) { th1[0] = colload(a.th, k + 1, 0); th1[1] = colload(a.th, k + 1, 1); }
            TF old0 = 0, old1 = 0;
            if (st_w) { old0 = a.wt[o_f]; old1 = a.wt[o_f + 1]; }
            const int oc = vorder(f, ks - 1, ke);
            const TF rho_c = p_rho[pl + 1], dzi_c = p_dzi[pl + 1], rdzhi_f = p_rdzhi[pl + 1];
            TF wx[12], E1R[2];
            row10(W1, wx); pair(E1, E1R);
            TF gt[2];
#pragma unroll
            for (int c = 0; c < 2; ++c)
            {
                // column of w around the centre of cell f: w[f-2..f+3] = wa(k-1), wb(k), W1(k+1), we(k+2), wc(k+3), wd(k+4)
                const TF wf = X10(wx, c);
                const TF ft_a = rho_c * vflux_col<TF>(oc, interp2(wf, we[c]), wa[c], wb[c], wf, we[c], wc[c], wd[c]);
                const TF ft_d = rho_c * (E1R[c] + visc) * (we[c] - wf) * dzi_c;
                gt[c] = TF(2.) * ft_d - ft_a;
            }
            if (st_w)
            {
                TF u4[6], u14[6], e0[8], e16[8], s0r[3], s1r[3];
                row4(U0, u4); row4(U1, u14); row6(E0, e0); row6(E1, e16);
                psum(e0, s0r); psum(e16, s1r);
                TF fx[3], dx_[3];
#pragma unroll
                for (int m = 0; m < 3; ++m)
                {
                    fx[m] = flux65(interp2(X4(u4, m), X4(u14, m)), X10(wx, m - 3), X10(wx, m - 2), X10(wx, m - 1), X10(wx, m), X10(wx, m + 1), X10(wx, m + 2));
                    const TF eviscx = q * (s0r[m] + s1r[m]) + visc;
                    dx_[m] = eviscx * ((X10(wx, m) - X10(wx, m - 1)) * dxi + (X4(u14, m) - X4(u4, m)) * dzhi_f);
                }
                TF wy[7][2], V0R[2], V0P[2], V1R[2], V1P[2], EM0[2], EP0[2], EM1[2], EP1[2];
                col7(W1, wy);
                pair(V0, V0R); pair(V0 + P, V0P); pair(V1, V1R); pair(V1 + P, V1P);
                pair(E0 - P, EM0); pair(E0 + P, EP0); pair(E1 - P, EM1); pair(E1 + P, EP1);
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    const TF fn = flux65(interp2(V0P[c], V1P[c]), wy[1][c], wy[2][c], wy[3][c], wy[4][c], wy[5][c], wy[6][c]);
                    const TF fs = flux65(interp2(V0R[c], V1R[c]), wy[0][c], wy[1][c], wy[2][c], wy[3][c], wy[4][c], wy[5][c]);
                    const TF eviscn = q * ((X6(e0, c) + X6(e16, c)) + (EP0[c] + EP1[c])) + visc;
                    const TF eviscs = q * ((EM0[c] + X6(e0, c)) + (EM1[c] + X6(e16, c))) + visc;
                    TF tw = -(fx[c + 1] - fx[c]) * dxi - (fn - fs) * dyi
                          + (dx_[c + 1] - dx_[c]) * dxi
                          + (eviscn * ((wy[4][c] - wy[3][c]) * dyi + (V1P[c] - V0P[c]) * dzhi_f)
                           - eviscs * ((wy[3][c] - wy[2][c]) * dyi + (V1R[c] - V0R[c]) * dzhi_f)) * dyi
                          + (gt[c] - gw[c]) * rdzhi_f;
                    if (BUOY) tw += p_gth[pl + 1] * (interp2(thk[c], th1[c]) - p_thh[pl + 1]);
                    a.wt[o_f + c] = (c == 0 ? old0 : old1) + tw;
                }
            }
            gw[0] = gt[0]; gw[1] = gt[1];
            wa[0] = wb[0]; wa[1] = wb[1]; wb[0] = X10(wx, 0); wb[1] = X10(wx, 1);
            we[0] = wc[0]; we[1] = wc[1]; wc[0] = wd[0]; wc[1] = wd[1]; wd[0] = wn0; wd[1] = wn1;
            thk[0] = th1[0]; thk[1] = th1[1];
        }
#undef X10
#undef X6
#undef X4
        BLOCK_FENCE();
#undef BLOCK_FENCE
        s0 = s1;
    }
}

} // namespace mhh
