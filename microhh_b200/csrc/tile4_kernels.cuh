// mhhb200 -- mom4: 2 x 2 register-blocked, warp-specialised TMA tile kernel for the fused tendencies
// (thermo.exec buoyancy + advec_2i5 + diff_smag2 of u, v, w and the first scalar).
//
// Successor of mom3_kernel (tile3_kernels.cuh), attacking the two ceilings its ncu profile shows (L1/shared-memory
// wavefronts 91 %, fp64 pipe 56 %):
//   * every thread owns TWO rows x TWO columns, so the three y-face fluxes of its two rows are computed once (3 instead
//     of 4), the y-direction column loads serve two rows (8 row reads for 2 rows instead of 7 per row) and the
//     viscosity / velocity rows shared by neighbouring rows are read once;
//   * every shared-memory row read is an aligned vector LDS; where the own pair sits at an odd column (fp64, see t4_hl)
//     the two vectors that cover it also deliver the x neighbours the stencils need, so no wavefront is wasted;
//   * a CTA covers 64 x (2*TYW) columns with TYW warps per component, so the halo amplification of the staged planes
//     drops from (3+6)/3 = 3.0 to (6+6)/6 = 2.0 rows per row.
// Warp roles as in mom3: groups of TYW warps compute u, v, w(+buoyancy) and the first scalar on the same TMA-staged
// planes (full/empty mbarrier per ring slot), one extra warp is the TMA producer.
//
// Arithmetic: flux form of reference src/advec_2i5.cxx:151-728, include/diff_kernels.h:144-484,
// src/thermo_dry.cxx:165-179 (same expressions as stencil_kernels.cuh; shared faces re-associate sums at the ulp level).
#pragma once
#include "tile2_kernels.cuh"
#include "tile3_kernels.cuh"

namespace mhh {

constexpr int T4_W = 64;                   // tile width: 2 columns per lane
// x halo.  The TMA box origin must be 16-byte aligned in global memory (measured: an odd fp64 x coordinate raises
// "illegal instruction"; negative even ones are fine and zero-filled): fp64 fields with igc = 3 start their interior at an
// odd element, so the box starts 3 columns to the left and the own pair sits at an ODD shared-memory column (every
// 128-bit LDS then starts one column early and also delivers a neighbour); fp32 fields with igc = 4 use a halo of 4
// and the own pair is an aligned 64-bit LDS.
constexpr int t4_hl(int elem) { return elem == 8 ? 3 : 4; }
constexpr int t4_px(int elem) { return T4_W + 2 * t4_hl(elem); }   // 70 / 72: plane pitch = TMA box width (multiples of 16 bytes)
constexpr int T4_H = 3;                    // y halo
constexpr int T4_RING = 4;
constexpr int t4_rows(int tyw) { return 2 * tyw + 2 * T4_H; }
constexpr int t4_plane(int tyw, int elem) { return (t4_px(elem) * t4_rows(tyw) * elem + 127) / 128 * 128 / elem; }
constexpr int t4_box_bytes(int tyw, int elem) { return t4_px(elem) * t4_rows(tyw) * elem; }

inline size_t mom4_smem(size_t elem, int kchunk, int tyw, int nsc)
{ return 128 + ((size_t)(4 + nsc) * T4_RING * t4_plane(tyw, (int)elem) + (size_t)8 * (kchunk + 3)) * elem + 128; }

template <typename TF, bool SURFACE, bool BUOY, int NSC, int TYW>
__global__ void __launch_bounds__(32 * ((3 + NSC) * TYW + 1), 1)
mom4_kernel(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_v,
            const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_e,
            const __grid_constant__ CUtensorMap tm_s,
            const __grid_constant__ CUtensorMap tm_ut, const __grid_constant__ CUtensorMap tm_vt,
            const __grid_constant__ CUtensorMap tm_wt, const __grid_constant__ CUtensorMap tm_st,
            const Tend3Args<TF> args, const GridDev<TF> g)
{
    typedef typename V2T<TF>::type V2;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sbase = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sbase);     // full[RING], empty[RING]
    TF* sm = reinterpret_cast<TF*>(sbase + 128);
    constexpr int RING = T4_RING, R = 2 * TYW;
    constexpr int PLANE = t4_plane(TYW, (int)sizeof(TF)), P = t4_px((int)sizeof(TF)), T4_HL = t4_hl((int)sizeof(TF));
    constexpr int NCW = (3 + NSC) * TYW, NT = 32 * (NCW + 1);
    constexpr bool ODD = (T4_HL & 1) != 0;          // own pair at an odd shared-memory column
    constexpr int NF = 4 + NSC;
    constexpr unsigned PLANE_BYTES = PLANE * sizeof(TF), BOX_BYTES = t4_box_bytes(TYW, (int)sizeof(TF));

    const MomArgs<TF>& a = args.m;
    const int warp = threadIdx.x >> 5, tx = threadIdx.x & 31;
    // NSC = 1: group = warp % 4, so each SM sub-partition runs ONE component's loop (one code path in its instruction cache)
    const int comp = NSC ? (warp & 3) : warp / TYW, tw = NSC ? (warp >> 2) : warp - (warp / TYW) * TYW;
    const int i = g.istart + blockIdx.x * T4_W + 2 * tx;
    const int j0 = g.jstart + blockIdx.y * R + 2 * tw;
    const int gi0 = g.istart + blockIdx.x * T4_W - T4_HL;          // 16-byte aligned (checked by the launcher)
    const int gj0 = g.jstart + blockIdx.y * R - T4_H;
    const bool xact = (i + 1 < g.iend);
    const bool act[2] = {xact && (j0 < g.jend), xact && (j0 + 1 < g.jend)};
    const int ks = g.kstart, ke = g.kend;
    const int kc0 = ks + blockIdx.z * args.kchunk;
    const int kc1 = min(ke, kc0 + args.kchunk);
    const long long jj = g.icells, kk = g.ijcells;
    const int ic = min(i, g.iend - 2);
    const long long ij[2] = {ic + min(j0, g.jend - 1) * jj, ic + min(j0 + 1, g.jend - 1) * jj};
    const int sidx = (2 * tw + T4_H) * P + T4_HL + 2 * tx;
    const TF dxi = g.dxi, dyi = g.dyi, visc = a.visc;
    const TF q = TF(0.25);

    const int k0 = kc0 - 1;
    TF* prof = sm + NF * RING * PLANE;
    const int nlev = args.kchunk + 3;
    TF* p_rho = prof; TF* p_rhoh = prof + nlev; TF* p_rdzi = prof + 2 * nlev; TF* p_rdzhi = prof + 3 * nlev;
    TF* p_dzhi = prof + 4 * nlev; TF* p_gth = prof + 5 * nlev; TF* p_thh = prof + 6 * nlev; TF* p_dzi = prof + 7 * nlev;
    for (int t = threadIdx.x; t < nlev; t += NT)
    {
        const int lev = min(max(k0 + t, 0), g.kcells - 1);
        const TF rho = g.rhoref[lev], rhoh = g.rhorefh[lev];
        p_rho[t] = rho; p_rhoh[t] = rhoh;
        p_rdzi[t] = g.dzi[lev] / rho;
        p_rdzhi[t] = g.dzhi[lev] / rhoh;
        p_dzhi[t] = g.dzhi[lev];
        p_gth[t] = BUOY ? TF(GRAV) / g.threfh[lev] : TF(0);
        p_thh[t] = BUOY ? g.threfh[lev] : TF(0);
        p_dzi[t] = g.dzi[lev];
    }
    const unsigned full0 = smem_u32(bars), empty0 = full0 + 8 * RING;
    const unsigned pl0 = smem_u32(sm);
    if (threadIdx.x == 0)
    {
        for (int s = 0; s < RING; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, NCW); }
        mbar_fence_init();
    }
    __syncthreads();            // the only block-wide barrier: roles split below

    // ================================================================== producer warp
    if (warp == NCW)
    {
        if (tx == 0)
        {
            for (int lev = k0; lev <= kc1; ++lev)
            {
                const int n = lev - k0, slot = n % RING, use = n / RING;
                if (use > 0) mbar_wait(empty0 + 8 * slot, (use - 1) & 1);       // every consumer warp released the slot
                const unsigned bar = full0 + 8 * slot;
                mbar_expect_tx(bar, NF * BOX_BYTES);
                tma_load_3d(pl0 + (0 * RING + slot) * PLANE_BYTES, &tm_u, bar, gi0, gj0, lev);
                tma_load_3d(pl0 + (1 * RING + slot) * PLANE_BYTES, &tm_v, bar, gi0, gj0, lev);
                tma_load_3d(pl0 + (2 * RING + slot) * PLANE_BYTES, &tm_w, bar, gi0, gj0, lev);
                tma_load_3d(pl0 + (3 * RING + slot) * PLANE_BYTES, &tm_e, bar, gi0, gj0, lev);
                if (NSC) tma_load_3d(pl0 + (4 * RING + slot) * PLANE_BYTES, &tm_s, bar, gi0, gj0, lev);
                if (args.prefetch)
                {
                    // what the consumers' per-thread loads touch a little later: leading window levels, tendencies
                    const int pu = lev + 2 + args.prefetch;
                    if (pu < g.kcells) { tma_prefetch_3d(&tm_u, gi0, gj0, pu); tma_prefetch_3d(&tm_v, gi0, gj0, pu); if (NSC) tma_prefetch_3d(&tm_s, gi0, gj0, pu); }
                    if (pu + 1 < g.kcells) tma_prefetch_3d(&tm_w, gi0, gj0, pu + 1);
                    const int pt = lev + args.prefetch - 1;
                    if (pt >= kc0 && pt < kc1)
                    {
                        tma_prefetch_3d(&tm_ut, gi0, gj0 + T4_H, pt); tma_prefetch_3d(&tm_vt, gi0, gj0 + T4_H, pt);
                        tma_prefetch_3d(&tm_wt, gi0, gj0 + T4_H, pt + 1);
                        if (NSC) tma_prefetch_3d(&tm_st, gi0, gj0 + T4_H, pt);
                    }
                }
            }
        }
        return;
    }

    // ================================================================== consumer warps
    auto colload = [&](const TF* __restrict__ fld, int lev, int r, int c) -> TF {
        return (lev >= 0 && lev < g.kcells) ? fld[ij[r] + c + (long long)lev * kk] : TF(0);
    };
    auto LD2 = [](const TF* p) -> V2 { return *reinterpret_cast<const V2*>(p); };
    // Aligned vector row readers around the own pair (x[0], x[1]).  ODD: vectors start at x[-3], x[-1], x[1], x[3];
    // otherwise at x[-4], x[-2], x[0], x[2], x[4].  x[n] lands at array index n + O? (the X? macros below).
    constexpr int OA = ODD ? 3 : 4, NA = ODD ? 4 : 5;      // rowA: x[-3..4] are used
    constexpr int OB = ODD ? 1 : 2, NB = ODD ? 2 : 3;      // rowB: x[-1..2] are used
    constexpr int OL = ODD ? 1 : 2;                        // rowL: x[-1..1] are used (two vectors)
    constexpr int OR_ = ODD ? 1 : 0;                       // rowR: x[0..2]  are used (two vectors)
    auto rowA = [&](const TF* p, TF (&x)[10]) {
#pragma unroll
        for (int n = 0; n < NA; ++n) { const V2 t = LD2(p + 2 * n - OA); x[2 * n] = t.x; x[2 * n + 1] = t.y; } };
    auto rowB = [&](const TF* p, TF (&x)[6]) {
#pragma unroll
        for (int n = 0; n < NB; ++n) { const V2 t = LD2(p + 2 * n - OB); x[2 * n] = t.x; x[2 * n + 1] = t.y; } };
    auto rowL = [&](const TF* p, TF (&x)[4]) {
        const V2 t0 = LD2(p - OL), t1 = LD2(p - OL + 2); x[0] = t0.x; x[1] = t0.y; x[2] = t1.x; x[3] = t1.y; };
    auto rowR = [&](const TF* p, TF (&x)[4]) {
        const V2 t0 = LD2(p - OR_), t1 = LD2(p - OR_ + 2); x[0] = t0.x; x[1] = t0.y; x[2] = t1.x; x[3] = t1.y; };
    auto pair = [&](const TF* p, TF (&x)[2]) {
        if (ODD) { x[0] = p[0]; x[1] = p[1]; } else { const V2 t = LD2(p); x[0] = t.x; x[1] = t.y; } };
    // pair sums along x of a viscosity row: S[m] = E[m-1] + E[m], m = 0..2
    auto psum = [&](const TF (&e)[6], TF (&sum)[3]) {
#pragma unroll
        for (int m = 0; m < 3; ++m) sum[m] = e[m - 1 + OB] + e[m + OB]; };
    auto plane = [&](int fld, int slot) -> const TF* { return sm + (fld * RING + slot) * PLANE + sidx; };
    auto acquire = [&](int k, int& s0, int& s1) {
        const int n = k - k0;
        s0 = n % RING; s1 = (n + 1) % RING;
        if (n == 0) mbar_wait(full0, 0);
        mbar_wait(full0 + 8 * s1, ((n + 1) / RING) & 1);
    };
    auto release = [&](int s0) { __syncwarp(); if (tx == 0) mbar_arrive(empty0 + 8 * s0); };

#define XA(a, n) a[(n) + OA]
#define XB(a, n) a[(n) + OB]
#define XL(a, n) a[(n) + OL]
#define XR(a, n) a[(n) + OR_]
    if (comp == 0)
    {
        // ------------------------------------------------------------------ u
        TF ua[2][2], ub[2][2], uc[2][2], ud[2][2], gu[2][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c)
            {
                ua[r][c] = colload(a.u, k0 - 2, r, c); ub[r][c] = colload(a.u, k0 - 1, r, c);
                uc[r][c] = colload(a.u, k0 + 2, r, c); ud[r][c] = colload(a.u, k0 + 3, r, c); gu[r][c] = TF(0);
            }
        for (int k = k0; k < kc1; ++k)
        {
            const int f = k + 1, pl = k - k0;
            const bool lev_st = (k >= kc0);
            TF un[2][2], old[2][2];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    un[r][c] = colload(a.u, k + 4, r, c);
                    old[r][c] = (lev_st && act[r]) ? a.ut[ij[r] + c + (long long)k * kk] : TF(0);
                }
            const TF rhoh_f = p_rhoh[pl + 1], rdzi_k = p_rdzi[pl], dzhi_f = p_dzhi[pl + 1];
            const int of = vorder(f, ks, ke);
            int s0, s1;
            acquire(k, s0, s1);
            const TF* __restrict__ U0 = plane(0, s0); const TF* __restrict__ U1 = plane(0, s1);
            const TF* __restrict__ V0 = plane(1, s0); const TF* __restrict__ W1 = plane(2, s1);
            const TF* __restrict__ E0 = plane(3, s0); const TF* __restrict__ E1 = plane(3, s1);
            TF ux[2][10], w6[2][4], e0[2][6], e1[2][6], s0r[2][3], s1r[2][3], U1R[2][2];
#pragma unroll
            for (int r = 0; r < 2; ++r)
            {
                rowA(U0 + r * P, ux[r]); rowL(W1 + r * P, w6[r]); rowB(E0 + r * P, e0[r]); rowB(E1 + r * P, e1[r]); pair(U1 + r * P, U1R[r]);
                psum(e0[r], s0r[r]); psum(e1[r], s1r[r]);
            }
            TF fx[2][3], dx_[2][3], gt[2][2];
#pragma unroll
            for (int r = 0; r < 2; ++r)
            {
#pragma unroll
                for (int m = 0; m < 3; ++m)
                {
                    fx[r][m] = flux65(interp2(XA(ux[r], m - 1), XA(ux[r], m)), XA(ux[r], m - 3), XA(ux[r], m - 2), XA(ux[r], m - 1), XA(ux[r], m), XA(ux[r], m + 1), XA(ux[r], m + 2));
                    dx_[r][m] = (XB(e0[r], m - 1) + visc) * (XA(ux[r], m) - XA(ux[r], m - 1)) * dxi;
                }
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    const TF uk = XA(ux[r], c), uk1 = U1R[r][c];
                    const TF ft_a = rhoh_f * vflux_col<TF>(of, interp2(XL(w6[r], c - 1), XL(w6[r], c)), ua[r][c], ub[r][c], uk, uk1, uc[r][c], ud[r][c]);
                    TF ft_d;
                    if (SURFACE && f == ks) ft_d = -rhoh_f * a.u_fluxbot[ij[r] + c];
                    else if (SURFACE && f == ke) ft_d = -rhoh_f * a.u_fluxtop[ij[r] + c];
                    else
                    {
                        const TF evisct = q * (s0r[r][c] + s1r[r][c]) + visc;
                        ft_d = rhoh_f * evisct * ((uk1 - uk) * dzhi_f + (XL(w6[r], c) - XL(w6[r], c - 1)) * dxi);
                    }
                    gt[r][c] = ft_d - ft_a;
                }
            }
            if (lev_st && act[0])
            {
                // y faces yf = 0..2: the south face of row yf.  u rows -3..4 around row 0, v rows 0..2, evisc pair sums of rows -1..2
                TF uy[8][2], v6[3][4], em[6], ep[6], sE[4][3];
#pragma unroll
                for (int d = -3; d <= 4; ++d)
                {
                    if (d == 0 || d == 1) { uy[d + 3][0] = XA(ux[d], 0); uy[d + 3][1] = XA(ux[d], 1); }
                    else pair(U0 + d * P, uy[d + 3]);
                }
#pragma unroll
                for (int yf = 0; yf < 3; ++yf) rowL(V0 + yf * P, v6[yf]);
                rowB(E0 - P, em); rowB(E0 + 2 * P, ep);
                psum(em, sE[0]); psum(ep, sE[3]);
#pragma unroll
                for (int m = 0; m < 3; ++m) { sE[1][m] = s0r[0][m]; sE[2][m] = s0r[1][m]; }
                TF F[3][2];          // advective - diffusive y-face flux
#pragma unroll
                for (int yf = 0; yf < 3; ++yf)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                    {
                        const TF fa = flux65(interp2(XL(v6[yf], c - 1), XL(v6[yf], c)), uy[yf][c], uy[yf + 1][c], uy[yf + 2][c], uy[yf + 3][c], uy[yf + 4][c], uy[yf + 5][c]);
                        const TF ev = q * (sE[yf][c] + sE[yf + 1][c]) + visc;
                        F[yf][c] = fa - ev * ((uy[yf + 3][c] - uy[yf + 2][c]) * dyi + (XL(v6[yf], c) - XL(v6[yf], c - 1)) * dxi);
                    }
#pragma unroll
                for (int r = 0; r < 2; ++r)
                    if (act[r])
                    {
                        const long long o_k = ij[r] + (long long)k * kk;
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                        {
                            const TF tu = -(fx[r][c + 1] - fx[r][c]) * dxi - (F[r + 1][c] - F[r][c]) * dyi
                                        + (dx_[r][c + 1] - dx_[r][c]) * TF(2.) * dxi + (gt[r][c] - gu[r][c]) * rdzi_k;
                            a.ut[o_k + c] = old[r][c] + tu;
                        }
                    }
            }
            release(s0);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    gu[r][c] = gt[r][c];
                    ua[r][c] = ub[r][c]; ub[r][c] = XA(ux[r], c);
                    uc[r][c] = ud[r][c]; ud[r][c] = un[r][c];
                }
        }
    }
    else if (comp == 1)
    {
        // ------------------------------------------------------------------ v
        TF va[2][2], vb[2][2], vc[2][2], vd[2][2], gv[2][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c)
            {
                va[r][c] = colload(a.v, k0 - 2, r, c); vb[r][c] = colload(a.v, k0 - 1, r, c);
                vc[r][c] = colload(a.v, k0 + 2, r, c); vd[r][c] = colload(a.v, k0 + 3, r, c); gv[r][c] = TF(0);
            }
        for (int k = k0; k < kc1; ++k)
        {
            const int f = k + 1, pl = k - k0;
            const bool lev_st = (k >= kc0);
            TF vn[2][2], old[2][2];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    vn[r][c] = colload(a.v, k + 4, r, c);
                    old[r][c] = (lev_st && act[r]) ? a.vt[ij[r] + c + (long long)k * kk] : TF(0);
                }
            const TF rhoh_f = p_rhoh[pl + 1], rdzi_k = p_rdzi[pl], dzhi_f = p_dzhi[pl + 1];
            const int of = vorder(f, ks, ke);
            int s0, s1;
            acquire(k, s0, s1);
            const TF* __restrict__ U0 = plane(0, s0);
            const TF* __restrict__ V0 = plane(1, s0); const TF* __restrict__ V1 = plane(1, s1);
            const TF* __restrict__ W1 = plane(2, s1);
            const TF* __restrict__ E0 = plane(3, s0); const TF* __restrict__ E1 = plane(3, s1);
            // rows -1, 0, 1 of evisc (level k: with x neighbours; level k+1: own pair) and of w
            TF vx[2][10], e0[3][6], V1R[2][2], W1P[3][2], E1P[3][2];
#pragma unroll
            for (int r = 0; r < 2; ++r) { rowA(V0 + r * P, vx[r]); pair(V1 + r * P, V1R[r]); }
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) { rowB(E0 + (rr - 1) * P, e0[rr]); pair(W1 + (rr - 1) * P, W1P[rr]); pair(E1 + (rr - 1) * P, E1P[rr]); }
            TF gt[2][2];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    const TF vk = XA(vx[r], c), vk1 = V1R[r][c];
                    const TF ft_a = rhoh_f * vflux_col<TF>(of, interp2(W1P[r][c], W1P[r + 1][c]), va[r][c], vb[r][c], vk, vk1, vc[r][c], vd[r][c]);
                    TF ft_d;
                    if (SURFACE && f == ks) ft_d = -rhoh_f * a.v_fluxbot[ij[r] + c];
                    else if (SURFACE && f == ke) ft_d = -rhoh_f * a.v_fluxtop[ij[r] + c];
                    else
                    {
                        const TF evisct = q * ((XB(e0[r], c) + XB(e0[r + 1], c)) + (E1P[r][c] + E1P[r + 1][c])) + visc;
                        ft_d = rhoh_f * evisct * ((vk1 - vk) * dzhi_f + (W1P[r + 1][c] - W1P[r][c]) * dyi);
                    }
                    gt[r][c] = ft_d - ft_a;
                }
            if (lev_st && act[0])
            {
                TF u4[3][4], vy[8][2], sE[3][3];
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) { rowR(U0 + (rr - 1) * P, u4[rr]); psum(e0[rr], sE[rr]); }
#pragma unroll
                for (int d = -3; d <= 4; ++d)
                {
                    if (d == 0 || d == 1) { vy[d + 3][0] = XA(vx[d], 0); vy[d + 3][1] = XA(vx[d], 1); }
                    else pair(V0 + d * P, vy[d + 3]);
                }
                TF fx[2][3];         // advective - diffusive x-face flux
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int m = 0; m < 3; ++m)
                    {
                        const TF fa = flux65(interp2(XR(u4[r], m), XR(u4[r + 1], m)), XA(vx[r], m - 3), XA(vx[r], m - 2), XA(vx[r], m - 1), XA(vx[r], m), XA(vx[r], m + 1), XA(vx[r], m + 2));
                        const TF eviscc = q * (sE[r][m] + sE[r + 1][m]) + visc;
                        fx[r][m] = fa - eviscc * ((XA(vx[r], m) - XA(vx[r], m - 1)) * dxi + (XR(u4[r + 1], m) - XR(u4[r], m)) * dyi);
                    }
                // y "faces" yf = 0..2 are the centres of cells j0-1+yf (evisc row yf-1 -> e0[yf])
                TF F[3][2];
#pragma unroll
                for (int yf = 0; yf < 3; ++yf)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                    {
                        const TF fa = flux65(interp2(vy[yf + 2][c], vy[yf + 3][c]), vy[yf][c], vy[yf + 1][c], vy[yf + 2][c], vy[yf + 3][c], vy[yf + 4][c], vy[yf + 5][c]);
                        F[yf][c] = fa - (XB(e0[yf], c) + visc) * (vy[yf + 3][c] - vy[yf + 2][c]) * dyi * TF(2.);
                    }
#pragma unroll
                for (int r = 0; r < 2; ++r)
                    if (act[r])
                    {
                        const long long o_k = ij[r] + (long long)k * kk;
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                        {
                            const TF tv = -(fx[r][c + 1] - fx[r][c]) * dxi - (F[r + 1][c] - F[r][c]) * dyi + (gt[r][c] - gv[r][c]) * rdzi_k;
                            a.vt[o_k + c] = old[r][c] + tv;
                        }
                    }
            }
            release(s0);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    gv[r][c] = gt[r][c];
                    va[r][c] = vb[r][c]; vb[r][c] = XA(vx[r], c);
                    vc[r][c] = vd[r][c]; vd[r][c] = vn[r][c];
                }
        }
    }
    else if (NSC == 1 && comp == 3)
    {
        // ------------------------------------------------------------------ scalar 0 (advec_s + diff_c)
        const ScalArgs<TF>& sa_ = args.sc;
        const TF h = TF(0.5), tPr_i = TF(1) / sa_.tPr, svisc = sa_.visc;
        TF sa[2][2], sb[2][2], sc[2][2], sd[2][2], gs[2][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c)
            {
                sa[r][c] = colload(sa_.s, k0 - 2, r, c); sb[r][c] = colload(sa_.s, k0 - 1, r, c);
                sc[r][c] = colload(sa_.s, k0 + 2, r, c); sd[r][c] = colload(sa_.s, k0 + 3, r, c); gs[r][c] = TF(0);
            }
        for (int k = k0; k < kc1; ++k)
        {
            const int f = k + 1, pl = k - k0;
            const bool lev_st = (k >= kc0);
            TF sn[2][2], old[2][2];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    sn[r][c] = colload(sa_.s, k + 4, r, c);
                    old[r][c] = (lev_st && act[r]) ? sa_.st[ij[r] + c + (long long)k * kk] : TF(0);
                }
            const TF rhoh_f = p_rhoh[pl + 1], rdzi_k = p_rdzi[pl], dzhi_f = p_dzhi[pl + 1];
            const int of = vorder(f, ks, ke);
            int s0, s1;
            acquire(k, s0, s1);
            const TF* __restrict__ U0 = plane(0, s0); const TF* __restrict__ V0 = plane(1, s0);
            const TF* __restrict__ W1 = plane(2, s1);
            const TF* __restrict__ E0 = plane(3, s0); const TF* __restrict__ E1 = plane(3, s1);
            const TF* __restrict__ S0 = plane(4, s0); const TF* __restrict__ S1 = plane(4, s1);
            TF sx[2][10], S1R[2][2], W1R[2][2], E1R[2][2], e0[2][6];
#pragma unroll
            for (int r = 0; r < 2; ++r)
            { rowA(S0 + r * P, sx[r]); pair(S1 + r * P, S1R[r]); pair(W1 + r * P, W1R[r]); pair(E1 + r * P, E1R[r]); rowB(E0 + r * P, e0[r]); }
            TF gt[2][2];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    const TF sk = XA(sx[r], c), sk1 = S1R[r][c];
                    const TF ft_a = rhoh_f * vflux_col<TF>(of, W1R[r][c], sa[r][c], sb[r][c], sk, sk1, sc[r][c], sd[r][c]);
                    TF ft_d;
                    if (SURFACE && f == ks) ft_d = -rhoh_f * sa_.fluxbot[ij[r] + c];
                    else if (SURFACE && f == ke) ft_d = -rhoh_f * sa_.fluxtop[ij[r] + c];
                    else
                    {
                        const TF evisct = h * (XB(e0[r], c) + E1R[r][c]) * tPr_i + svisc;
                        ft_d = rhoh_f * evisct * (sk1 - sk) * dzhi_f;
                    }
                    gt[r][c] = ft_d - ft_a;
                }
            if (lev_st && act[0])
            {
                TF u4[2][4], sy[8][2], V0P[3][2], EP[4][2];
#pragma unroll
                for (int r = 0; r < 2; ++r) rowR(U0 + r * P, u4[r]);
#pragma unroll
                for (int d = -3; d <= 4; ++d)
                {
                    if (d == 0 || d == 1) { sy[d + 3][0] = XA(sx[d], 0); sy[d + 3][1] = XA(sx[d], 1); }
                    else pair(S0 + d * P, sy[d + 3]);
                }
#pragma unroll
                for (int yf = 0; yf < 3; ++yf) pair(V0 + yf * P, V0P[yf]);
                pair(E0 - P, EP[0]); pair(E0 + 2 * P, EP[3]);
#pragma unroll
                for (int c = 0; c < 2; ++c) { EP[1][c] = XB(e0[0], c); EP[2][c] = XB(e0[1], c); }
                TF fx[2][3];         // advective x-face flux; diffusive one kept apart (different metric factor)
                TF dx_[2][3];
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int m = 0; m < 3; ++m)
                    {
                        fx[r][m] = flux65(XR(u4[r], m), XA(sx[r], m - 3), XA(sx[r], m - 2), XA(sx[r], m - 1), XA(sx[r], m), XA(sx[r], m + 1), XA(sx[r], m + 2));
                        const TF eviscx = h * (XB(e0[r], m - 1) + XB(e0[r], m)) * tPr_i + svisc;
                        dx_[r][m] = eviscx * (XA(sx[r], m) - XA(sx[r], m - 1));
                    }
                TF F[3][2], DY[3][2];
#pragma unroll
                for (int yf = 0; yf < 3; ++yf)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                    {
                        F[yf][c] = flux65(V0P[yf][c], sy[yf][c], sy[yf + 1][c], sy[yf + 2][c], sy[yf + 3][c], sy[yf + 4][c], sy[yf + 5][c]);
                        const TF evy = h * (EP[yf][c] + EP[yf + 1][c]) * tPr_i + svisc;
                        DY[yf][c] = evy * (sy[yf + 3][c] - sy[yf + 2][c]);
                    }
#pragma unroll
                for (int r = 0; r < 2; ++r)
                    if (act[r])
                    {
                        const long long o_k = ij[r] + (long long)k * kk;
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                        {
                            const TF d = (dx_[r][c + 1] - dx_[r][c]) * sa_.dxidxi + (DY[r + 1][c] - DY[r][c]) * sa_.dyidyi;
                            const TF ts = -(fx[r][c + 1] - fx[r][c]) * dxi - (F[r + 1][c] - F[r][c]) * dyi + d + (gt[r][c] - gs[r][c]) * rdzi_k;
                            sa_.st[o_k + c] = old[r][c] + ts;
                        }
                    }
            }
            release(s0);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    gs[r][c] = gt[r][c];
                    sa[r][c] = sb[r][c]; sb[r][c] = XA(sx[r], c);
                    sc[r][c] = sd[r][c]; sd[r][c] = sn[r][c];
                }
        }
    }
    else
    {
        // ------------------------------------------------------------------ w at face f = k+1 (+ buoyancy)
        TF wa[2][2], wb[2][2], we[2][2], wc[2][2], wd[2][2], gw[2][2], thc[2][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c)
            {
                wa[r][c] = colload(a.w, k0 - 1, r, c); wb[r][c] = colload(a.w, k0, r, c); we[r][c] = colload(a.w, k0 + 2, r, c);
                wc[r][c] = colload(a.w, k0 + 3, r, c); wd[r][c] = colload(a.w, k0 + 4, r, c); gw[r][c] = TF(0);
                thc[r][c] = (BUOY && !NSC) ? colload(a.th, k0, r, c) : TF(0);
            }
        for (int k = k0; k < kc1; ++k)
        {
            const int f = k + 1, pl = k - k0;
            const bool lev_st = (k >= kc0) && f < ke;
            TF wn[2][2], old[2][2], thn[2][2];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    wn[r][c] = colload(a.w, k + 5, r, c);
                    thn[r][c] = (BUOY && !NSC) ? colload(a.th, k + 1, r, c) : TF(0);     // th[k+1] straight from global when the scalar is not staged
                    old[r][c] = (lev_st && act[r]) ? a.wt[ij[r] + c + (long long)f * kk] : TF(0);
                }
            const int oc = vorder(f, ks - 1, ke);
            const TF rho_c = p_rho[pl + 1], dzi_c = p_dzi[pl + 1], rdzhi_f = p_rdzhi[pl + 1], dzhi_f = p_dzhi[pl + 1];
            int s0, s1;
            acquire(k, s0, s1);
            const TF* __restrict__ U0 = plane(0, s0); const TF* __restrict__ U1 = plane(0, s1);
            const TF* __restrict__ V0 = plane(1, s0); const TF* __restrict__ V1 = plane(1, s1);
            const TF* __restrict__ W1 = plane(2, s1);
            const TF* __restrict__ E0 = plane(3, s0); const TF* __restrict__ E1 = plane(3, s1);
            TF wx[2][10], E1R[2][2], thk[2][2], th1[2][2];
#pragma unroll
            for (int r = 0; r < 2; ++r)
            {
                rowA(W1 + r * P, wx[r]); pair(E1 + r * P, E1R[r]);
                if (BUOY && NSC) { pair(plane(4, s0) + r * P, thk[r]); pair(plane(4, s1) + r * P, th1[r]); }       // th rides in the scalar planes
                else
                {
#pragma unroll
                    for (int c = 0; c < 2; ++c) { thk[r][c] = thc[r][c]; th1[r][c] = thn[r][c]; thc[r][c] = thn[r][c]; }
                }
            }
            TF gt[2][2];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    const TF wf = XA(wx[r], c);
                    const TF ft_a = rho_c * vflux_col<TF>(oc, interp2(wf, we[r][c]), wa[r][c], wb[r][c], wf, we[r][c], wc[r][c], wd[r][c]);
                    const TF ft_d = rho_c * (E1R[r][c] + visc) * (we[r][c] - wf) * dzi_c;
                    gt[r][c] = TF(2.) * ft_d - ft_a;
                }
            if (lev_st && act[0])
            {
                TF u4[2][4], u14[2][4], e0[2][6], e16[2][6], s0r[2][3], s1r[2][3];
#pragma unroll
                for (int r = 0; r < 2; ++r)
                {
                    rowR(U0 + r * P, u4[r]); rowR(U1 + r * P, u14[r]); rowB(E0 + r * P, e0[r]); rowB(E1 + r * P, e16[r]);
                    psum(e0[r], s0r[r]); psum(e16[r], s1r[r]);
                }
                TF fx[2][3];         // advective - diffusive x-face flux
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int m = 0; m < 3; ++m)
                    {
                        const TF fa = flux65(interp2(XR(u4[r], m), XR(u14[r], m)), XA(wx[r], m - 3), XA(wx[r], m - 2), XA(wx[r], m - 1), XA(wx[r], m), XA(wx[r], m + 1), XA(wx[r], m + 2));
                        const TF eviscx = q * (s0r[r][m] + s1r[r][m]) + visc;
                        fx[r][m] = fa - eviscx * ((XA(wx[r], m) - XA(wx[r], m - 1)) * dxi + (XR(u14[r], m) - XR(u4[r], m)) * dzhi_f);
                    }
                TF wy[8][2], V0P[3][2], V1P[3][2], EP0[4][2], EP1[4][2];
#pragma unroll
                for (int d = -3; d <= 4; ++d)
                {
                    if (d == 0 || d == 1) { wy[d + 3][0] = XA(wx[d], 0); wy[d + 3][1] = XA(wx[d], 1); }
                    else pair(W1 + d * P, wy[d + 3]);
                }
#pragma unroll
                for (int yf = 0; yf < 3; ++yf) { pair(V0 + yf * P, V0P[yf]); pair(V1 + yf * P, V1P[yf]); }
                pair(E0 - P, EP0[0]); pair(E0 + 2 * P, EP0[3]); pair(E1 - P, EP1[0]); pair(E1 + 2 * P, EP1[3]);
#pragma unroll
                for (int c = 0; c < 2; ++c)
                { EP0[1][c] = XB(e0[0], c); EP0[2][c] = XB(e0[1], c); EP1[1][c] = XB(e16[0], c); EP1[2][c] = XB(e16[1], c); }
                TF F[3][2];
#pragma unroll
                for (int yf = 0; yf < 3; ++yf)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                    {
                        const TF fa = flux65(interp2(V0P[yf][c], V1P[yf][c]), wy[yf][c], wy[yf + 1][c], wy[yf + 2][c], wy[yf + 3][c], wy[yf + 4][c], wy[yf + 5][c]);
                        const TF evy = q * ((EP0[yf][c] + EP0[yf + 1][c]) + (EP1[yf][c] + EP1[yf + 1][c])) + visc;
                        F[yf][c] = fa - evy * ((wy[yf + 3][c] - wy[yf + 2][c]) * dyi + (V1P[yf][c] - V0P[yf][c]) * dzhi_f);
                    }
#pragma unroll
                for (int r = 0; r < 2; ++r)
                    if (act[r])
                    {
                        const long long o_f = ij[r] + (long long)f * kk;
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                        {
                            TF tw_ = -(fx[r][c + 1] - fx[r][c]) * dxi - (F[r + 1][c] - F[r][c]) * dyi + (gt[r][c] - gw[r][c]) * rdzhi_f;
                            if (BUOY) tw_ += p_gth[pl + 1] * (interp2(thk[r][c], th1[r][c]) - p_thh[pl + 1]);
                            a.wt[o_f + c] = old[r][c] + tw_;
                        }
                    }
            }
            release(s0);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    gw[r][c] = gt[r][c];
                    wa[r][c] = wb[r][c]; wb[r][c] = XA(wx[r], c);
                    we[r][c] = wc[r][c]; wc[r][c] = wd[r][c]; wd[r][c] = wn[r][c];
                }
        }
    }
#undef XA
#undef XB
#undef XL
#undef XR
}

} // namespace mhh
