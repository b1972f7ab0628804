// mhhb200 -- warp-specialised version of the TMA tile kernel for the fused momentum tendencies.
//
// Same tile geometry, shared-memory planes and arithmetic as mom2_kernel (tile2_kernels.cuh), but the
// three velocity components are computed by three different groups of warps of the same CTA, all
// reading the same TMA-staged planes:
//     warps [0, TY)      u tendency   (row ty = warp)
//     warps [TY, 2TY)    v tendency
//     warps [2TY, 3TY)   w tendency (+ buoyancy)
//     warps [3TY, 4TY)   first scalar (advec_s + diff_c), optional
//     last warp          producer: one lane issues the TMA loads / L2 prefetches
// A thread therefore carries the register window and the carried vertical flux of ONE component only
// (about a third of the persistent registers), which lets 3x more warps be resident per tile; the
// per-level block barrier is replaced by full/empty mbarriers per ring slot, so warps drift apart by up
// to RING-2 levels instead of all waiting for the slowest one.  One extra warp is the TMA producer.
// (Registers are allocated per SM sub-partition: with 17 warps one sub-partition holds 5 and everybody's
// budget drops from 128 to 96 registers, so warp counts of 4n+1 are avoided where registers are tight.)
#pragma once
#include "tile2_kernels.cuh"

namespace mhh {

__device__ __forceinline__ void mbar_arrive(unsigned bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory"); }

constexpr int T3_RING = 4;

inline size_t mom3_smem(size_t elem, int kchunk, int ty, int nsc, int hl)
{ return 128 + ((size_t)(4 + nsc) * T3_RING * t2_plane(ty, (int)elem, hl) + (size_t)8 * (kchunk + 3)) * elem + 128; }
// consumer warps / threads of a CTA: (3 + NSC) groups of TY rows, or TY rows of the scalar alone (SONLY)
constexpr int mom3_ncw(int nsc, int ty, bool sonly) { return sonly ? ty : (3 + nsc) * ty; }
// rows of the scalar-only variant: fp32 12 (13 warps like the fused kernel, two CTAs per SM), fp64 8 (five fields x four ring
// slots x the plane must fit the 227 KB of shared memory beside the profiles)
template <typename TF> constexpr int t3_sonly_ty() { return sizeof(TF) == 4 ? 12 : 8; }

// the first prognostic scalar rides along as a fourth warp group (NSC = 1): advec_s + diff_c on the same planes
template <typename TF>
struct Tend3Args
{
    MomArgs<TF> m;
    ScalArgs<TF> sc;
    int kchunk;
    int prefetch;
};

// resident CTAs per SM the kernel is compiled for: fp64 needs its 128 registers (one CTA); the fp32 variant uses 86, its planes
// are half as large, and two CTAs (26 warps) hide more latency than the few spills of a 72-register cap cost
template <typename TF> constexpr int mom3_min_blocks() { return sizeof(TF) == 4 ? 2 : 1; }

// ADV2: the advective fluxes of Advec_2 (src/advec_2.cxx:48-202: velocity * interp2 on every face) instead of Advec_2i5's --
// cases/drycblles as shipped (swadvec = 2 with smag2) on the same staged planes.
// SONLY (with NSC = 1): every consumer warp is a scalar warp -- advec_s + diff_c of ONE further prognostic scalar on the same
// staged planes (u, v, w, the scalar's eddy viscosity, the scalar); the momentum tendencies are not touched.
template <typename TF, bool SURFACE, bool BUOY, int NSC, int TY, int HL, bool ADV2 = false, bool SONLY = false>
__global__ void __launch_bounds__(32 * (mom3_ncw(NSC, TY, SONLY) + 1), mom3_min_blocks<TF>())
mom3_kernel(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_v,
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
    constexpr int RING = T3_RING;
    constexpr int PLANE = t2_plane(TY, (int)sizeof(TF), HL), P = t2_px(HL), T2_HL = HL;
    constexpr int NCW = mom3_ncw(NSC, TY, SONLY), NT = 32 * (NCW + 1);
    static_assert(!SONLY || (NSC == 1 && !BUOY), "the scalar-only variant is the scalar group alone");
    constexpr bool ODD = (HL & 1) != 0;             // own pair at an odd shared-memory column (fp64 with igc = 3)
    constexpr int NF = 4 + NSC;
    constexpr unsigned PLANE_BYTES = PLANE * sizeof(TF), BOX_BYTES = t2_box_bytes(TY, (int)sizeof(TF), HL);

    const MomArgs<TF>& a = args.m;
    const int warp = threadIdx.x >> 5, tx = threadIdx.x & 31;
    // NSC = 1: four groups, group = warp % 4.  Warps are dealt round-robin to the four SM sub-partitions, so each
    // sub-partition then runs ONE component's loop and its instruction cache holds one code path instead of four
    // (measured: 14 % of the stall samples were instruction fetches with group = warp / TY).
    const int comp = SONLY ? 3 : (NSC ? (warp & 3) : warp / TY), ty = SONLY ? warp : (NSC ? (warp >> 2) : warp - (warp / TY) * TY);
    const int i = g.istart + blockIdx.x * T2_W + 2 * tx;
    const int j = g.jstart + blockIdx.y * TY + ty;
    const int gi0 = g.istart + blockIdx.x * T2_W - T2_HL;
    const int gj0 = g.jstart + blockIdx.y * TY - T2_H;
    const bool active = (i + 1 < g.iend) && (j < g.jend);
    const int ks = g.kstart, ke = g.kend;
    const int kc0 = ks + blockIdx.z * args.kchunk;
    const int kc1 = min(ke, kc0 + args.kchunk);
    const long long jj = g.icells, kk = g.ijcells;
    const int ic = min(i, g.iend - 2), jc = min(j, g.jend - 1);
    const long long ij = ic + jc * jj;
    const int sidx = (ty + T2_H) * P + T2_HL + 2 * tx;
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
    // A dedicated warp keeps the TMA issue (a dozen slow uniform-datapath instructions per level) off the critical
    // path of the compute warps: letting the last warp that releases a slot refill it cost +25 % kernel time.
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
                        if (!SONLY)
                        {
                            tma_prefetch_3d(&tm_ut, gi0, gj0 + T2_H, pt); tma_prefetch_3d(&tm_vt, gi0, gj0 + T2_H, pt);
                            tma_prefetch_3d(&tm_wt, gi0, gj0 + T2_H, pt + 1);
                        }
                        if (NSC) tma_prefetch_3d(&tm_st, gi0, gj0 + T2_H, pt);
                    }
                }
            }
        }
        return;
    }

    // ================================================================== consumer warps
    // (Measured and rejected, profiles/r02/ab_mom3_sum2_*.json: passing face velocities as the SUM of the two neighbours and
    // folding the exact factor 1/2 into the metric saves 13.5 multiplications per point bit for bit, yet ran 2.5 % slower in
    // fp64 -- 17.2 / 17.6 vs 17.8 / 17.9 ms per step at 512^3: the register allocation under the 128-register cap got worse.)
    auto hflux = [](const TF vel, const TF x0, const TF x1, const TF x2, const TF x3, const TF x4, const TF x5) -> TF {
        return ADV2 ? flux2(vel, x2, x3) : flux65(vel, x0, x1, x2, x3, x4, x5); };
    auto vfl = [](const int order, const TF vel, const TF c0, const TF c1, const TF c2, const TF c3, const TF c4, const TF c5) -> TF {
        return vflux_col<TF>(ADV2 ? min(order, 2) : order, vel, c0, c1, c2, c3, c4, c5); };
    auto colload = [&](const TF* __restrict__ fld, int lev, int c) -> TF {
        return (lev >= 0 && lev < g.kcells) ? fld[ij + c + (long long)lev * kk] : TF(0);
    };
    auto LD2 = [](const TF* p) -> V2 { return *reinterpret_cast<const V2*>(p); };
    // the own pair in GLOBAL memory: one vector access when it is aligned (HL = 4), two scalar ones otherwise
    auto gload2 = [&](const TF* __restrict__ p, TF& x0, TF& x1) {
        if (ODD) { x0 = p[0]; x1 = p[1]; } else { const V2 t = *reinterpret_cast<const V2*>(p); x0 = t.x; x1 = t.y; } };
    auto gstore2 = [&](TF* __restrict__ p, const TF x0, const TF x1) {
        if (ODD) { p[0] = x0; p[1] = x1; } else { V2 t; t.x = x0; t.y = x1; *reinterpret_cast<V2*>(p) = t; } };
    auto colload2 = [&](const TF* __restrict__ fld, int lev, TF& x0, TF& x1) {
        if (lev >= 0 && lev < g.kcells) gload2(fld + ij + (long long)lev * kk, x0, x1); else { x0 = TF(0); x1 = TF(0); } };
    // Aligned vector row readers around the own pair (x[0], x[1]).  ODD: vectors start at x[-5], x[-3], x[-1], x[1] ...;
    // otherwise at x[-4], x[-2], x[0], x[2] ...  The value x[n] lands at array index n + O? (macros X10 / X6 / X4 below).
    constexpr int O10 = ODD ? 5 : 4, N10 = ODD ? 6 : 5;          // row10: x[-3..4] are used
    constexpr int O3 = ODD ? 1 : 2, N3 = ODD ? 2 : 3;            // row3:  x[-1..2] are used
    auto row10 = [&](const TF* p, TF (&x)[12]) {
#pragma unroll
        for (int n = 0; n < N10; ++n) { const V2 t = LD2(p + 2 * n - O10); x[2 * n] = t.x; x[2 * n + 1] = t.y; } };
    auto row3 = [&](const TF* p, TF (&x)[6]) {
#pragma unroll
        for (int n = 0; n < N3; ++n) { const V2 t = LD2(p + 2 * n - O3); x[2 * n] = t.x; x[2 * n + 1] = t.y; } };
    auto col7 = [&](const TF* p, TF (&y)[7][2]) {
#pragma unroll
        for (int d = -3; d <= 3; ++d)
        {
            if (ODD) { y[d + 3][0] = p[d * P]; y[d + 3][1] = p[d * P + 1]; }
            else { const V2 t = LD2(p + d * P); y[d + 3][0] = t.x; y[d + 3][1] = t.y; }
        } };
    auto pair = [&](const TF* p, TF (&x)[2]) {
        if (ODD) { x[0] = p[0]; x[1] = p[1]; } else { const V2 t = LD2(p); x[0] = t.x; x[1] = t.y; } };
    auto psum = [&](const TF (&e)[6], TF (&sum)[3]) {
#pragma unroll
        for (int m = 0; m < 3; ++m) sum[m] = e[m - 1 + O3] + e[m + O3]; };
    auto plane = [&](int fld, int slot) -> const TF* { return sm + (fld * RING + slot) * PLANE + sidx; };
    // wait for plane (k+1), hand back the slots of planes k and k+1
    auto acquire = [&](int k, int& s0, int& s1) {
        const int n = k - k0;
        s0 = n % RING; s1 = (n + 1) % RING;
        if (n == 0) mbar_wait(full0, 0);
        mbar_wait(full0 + 8 * s1, ((n + 1) / RING) & 1);
    };
    auto release = [&](int k, int s0) { (void)k; __syncwarp(); if (tx == 0) mbar_arrive(empty0 + 8 * s0); };

#define X10(a, n) a[(n) + O10]
#define X6(a, n) a[(n) + O3]
#define X4(a, n) a[(n) + O3]
    if (comp == 0)
    {
        // ------------------------------------------------------------------ u
        TF ua[2], ub[2], uc[2], ud[2], gu[2] = {0, 0};
#pragma unroll
        for (int c = 0; c < 2; ++c)
        { ua[c] = colload(a.u, k0 - 2, c); ub[c] = colload(a.u, k0 - 1, c); uc[c] = colload(a.u, k0 + 2, c); ud[c] = colload(a.u, k0 + 3, c); }
        for (int k = k0; k < kc1; ++k)
        {
            const int f = k + 1, pl = k - k0;
            const bool st = (k >= kc0) && active;
            const long long o_k = ij + (long long)k * kk;
            TF un0, un1; colload2(a.u, k + 4, un0, un1);
            TF old0 = 0, old1 = 0;
            if (st) gload2(a.ut + o_k, old0, old1);
            const TF rhoh_f = p_rhoh[pl + 1], rdzi_k = p_rdzi[pl], dzhi_f = p_dzhi[pl + 1];
            const int of = vorder(f, ks, ke);
            int s0, s1;
            acquire(k, s0, s1);
            const TF* __restrict__ U0 = plane(0, s0); const TF* __restrict__ U1 = plane(0, s1);
            const TF* __restrict__ V0 = plane(1, s0); const TF* __restrict__ W1 = plane(2, s1);
            const TF* __restrict__ E0 = plane(3, s0); const TF* __restrict__ E1 = plane(3, s1);
            TF ux[12], w6[6], e0[6], e1[6], s0r[3], s1r[3], U1R[2];
            row10(U0, ux); row3(W1, w6); row3(E0, e0); row3(E1, e1); pair(U1, U1R);
            psum(e0, s0r); psum(e1, s1r);
            TF fx[3], dx_[3];
#pragma unroll
            for (int m = 0; m < 3; ++m)
            {
                fx[m] = hflux(interp2(X10(ux, m - 1), X10(ux, m)), X10(ux, m - 3), X10(ux, m - 2), X10(ux, m - 1), X10(ux, m), X10(ux, m + 1), X10(ux, m + 2));
                dx_[m] = (X6(e0, m - 1) + visc) * (X10(ux, m) - X10(ux, m - 1)) * dxi;
            }
            TF gt[2];
#pragma unroll
            for (int c = 0; c < 2; ++c)
            {
                const TF uk = X10(ux, c), uk1 = U1R[c];
                const TF ft_a = rhoh_f * vfl(of, interp2(X6(w6, c - 1), X6(w6, c)), ua[c], ub[c], uk, uk1, uc[c], ud[c]);
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
            if (st)
            {
                TF uy[7][2], v6[6], vp6[6], em[6], ep[6], s0m[3], s0p[3];
                col7(U0, uy); row3(V0, v6); row3(V0 + P, vp6); row3(E0 - P, em); row3(E0 + P, ep);
                psum(em, s0m); psum(ep, s0p);
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    const TF fn = hflux(interp2(X6(vp6, c - 1), X6(vp6, c)), uy[1][c], uy[2][c], uy[3][c], uy[4][c], uy[5][c], uy[6][c]);
                    const TF fs = hflux(interp2(X6(v6, c - 1), X6(v6, c)), uy[0][c], uy[1][c], uy[2][c], uy[3][c], uy[4][c], uy[5][c]);
                    const TF eviscn = q * (s0r[c] + s0p[c]) + visc;
                    const TF eviscs = q * (s0m[c] + s0r[c]) + visc;
                    const TF d = (dx_[c + 1] - dx_[c]) * TF(2.) * dxi
                               + (eviscn * ((uy[4][c] - uy[3][c]) * dyi + (X6(vp6, c) - X6(vp6, c - 1)) * dxi)
                                - eviscs * ((uy[3][c] - uy[2][c]) * dyi + (X6(v6, c) - X6(v6, c - 1)) * dxi)) * dyi;
                    const TF tu = -(fx[c + 1] - fx[c]) * dxi - (fn - fs) * dyi + d + (gt[c] - gu[c]) * rdzi_k;
                    if (c == 0) old0 += tu; else old1 += tu;
                }
                gstore2(a.ut + o_k, old0, old1);
            }
            release(k, s0);
            gu[0] = gt[0]; gu[1] = gt[1];
            ua[0] = ub[0]; ua[1] = ub[1]; ub[0] = X10(ux, 0); ub[1] = X10(ux, 1);
            uc[0] = ud[0]; uc[1] = ud[1]; ud[0] = un0; ud[1] = un1;
        }
    }
    else if (comp == 1)
    {
        // ------------------------------------------------------------------ v
        TF va[2], vb[2], vc[2], vd[2], gv[2] = {0, 0};
#pragma unroll
        for (int c = 0; c < 2; ++c)
        { va[c] = colload(a.v, k0 - 2, c); vb[c] = colload(a.v, k0 - 1, c); vc[c] = colload(a.v, k0 + 2, c); vd[c] = colload(a.v, k0 + 3, c); }
        for (int k = k0; k < kc1; ++k)
        {
            const int f = k + 1, pl = k - k0;
            const bool st = (k >= kc0) && active;
            const long long o_k = ij + (long long)k * kk;
            TF vn0, vn1; colload2(a.v, k + 4, vn0, vn1);
            TF old0 = 0, old1 = 0;
            if (st) gload2(a.vt + o_k, old0, old1);
            const TF rhoh_f = p_rhoh[pl + 1], rdzi_k = p_rdzi[pl], dzhi_f = p_dzhi[pl + 1];
            const int of = vorder(f, ks, ke);
            int s0, s1;
            acquire(k, s0, s1);
            const TF* __restrict__ U0 = plane(0, s0);
            const TF* __restrict__ V0 = plane(1, s0); const TF* __restrict__ V1 = plane(1, s1);
            const TF* __restrict__ W1 = plane(2, s1);
            const TF* __restrict__ E0 = plane(3, s0); const TF* __restrict__ E1 = plane(3, s1);
            TF vx[12], e0[6], em[6], V1R[2], W1R[2], W1M[2], E1R[2], E1M[2];
            row10(V0, vx); row3(E0, e0); row3(E0 - P, em);
            pair(V1, V1R); pair(W1, W1R); pair(W1 - P, W1M); pair(E1, E1R); pair(E1 - P, E1M);
            TF gt[2];
#pragma unroll
            for (int c = 0; c < 2; ++c)
            {
                const TF vk = X10(vx, c), vk1 = V1R[c];
                const TF ft_a = rhoh_f * vfl(of, interp2(W1M[c], W1R[c]), va[c], vb[c], vk, vk1, vc[c], vd[c]);
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
            if (st)
            {
                TF u4[6], um4[6], vy[7][2], s0r[3], s0m[3];
                row3(U0, u4); row3(U0 - P, um4); col7(V0, vy);
                psum(e0, s0r); psum(em, s0m);
                TF fx[3], dx_[3];
#pragma unroll
                for (int m = 0; m < 3; ++m)
                {
                    fx[m] = hflux(interp2(X4(um4, m), X4(u4, m)), X10(vx, m - 3), X10(vx, m - 2), X10(vx, m - 1), X10(vx, m), X10(vx, m + 1), X10(vx, m + 2));
                    const TF eviscc = q * (s0m[m] + s0r[m]) + visc;
                    dx_[m] = eviscc * ((X10(vx, m) - X10(vx, m - 1)) * dxi + (X4(u4, m) - X4(um4, m)) * dyi);
                }
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    const TF fn = hflux(interp2(vy[3][c], vy[4][c]), vy[1][c], vy[2][c], vy[3][c], vy[4][c], vy[5][c], vy[6][c]);
                    const TF fs = hflux(interp2(vy[2][c], vy[3][c]), vy[0][c], vy[1][c], vy[2][c], vy[3][c], vy[4][c], vy[5][c]);
                    const TF d = (dx_[c + 1] - dx_[c]) * dxi
                               + ((X6(e0, c) + visc) * (vy[4][c] - vy[3][c]) * dyi - (X6(em, c) + visc) * (vy[3][c] - vy[2][c]) * dyi) * TF(2.) * dyi;
                    const TF tv = -(fx[c + 1] - fx[c]) * dxi - (fn - fs) * dyi + d + (gt[c] - gv[c]) * rdzi_k;
                    if (c == 0) old0 += tv; else old1 += tv;
                }
                gstore2(a.vt + o_k, old0, old1);
            }
            release(k, s0);
            gv[0] = gt[0]; gv[1] = gt[1];
            va[0] = vb[0]; va[1] = vb[1]; vb[0] = X10(vx, 0); vb[1] = X10(vx, 1);
            vc[0] = vd[0]; vc[1] = vd[1]; vd[0] = vn0; vd[1] = vn1;
        }
    }
    else if (NSC == 1 && comp == 3)
    {
        // ------------------------------------------------------------------ scalar 0 (advec_s + diff_c)
        const ScalArgs<TF>& sa_ = args.sc;
        const TF h = TF(0.5), tPr_i = TF(1) / sa_.tPr, svisc = sa_.visc;
        TF sa[2], sb[2], sc[2], sd[2], gs[2] = {0, 0};
#pragma unroll
        for (int c = 0; c < 2; ++c)
        { sa[c] = colload(sa_.s, k0 - 2, c); sb[c] = colload(sa_.s, k0 - 1, c); sc[c] = colload(sa_.s, k0 + 2, c); sd[c] = colload(sa_.s, k0 + 3, c); }
        for (int k = k0; k < kc1; ++k)
        {
            const int f = k + 1, pl = k - k0;
            const bool st = (k >= kc0) && active;
            const long long o_k = ij + (long long)k * kk;
            TF sn0, sn1; colload2(sa_.s, k + 4, sn0, sn1);
            TF old0 = 0, old1 = 0;
            if (st) gload2(sa_.st + o_k, old0, old1);
            const TF rhoh_f = p_rhoh[pl + 1], rdzi_k = p_rdzi[pl], dzhi_f = p_dzhi[pl + 1];
            const int of = vorder(f, ks, ke);
            int s0, s1;
            acquire(k, s0, s1);
            const TF* __restrict__ U0 = plane(0, s0); const TF* __restrict__ V0 = plane(1, s0);
            const TF* __restrict__ W1 = plane(2, s1);
            const TF* __restrict__ E0 = plane(3, s0); const TF* __restrict__ E1 = plane(3, s1);
            const TF* __restrict__ S0 = plane(4, s0); const TF* __restrict__ S1 = plane(4, s1);
            TF sx[12], S1R[2], W1R[2], E1R[2], e0[6];
            row10(S0, sx); pair(S1, S1R); pair(W1, W1R); pair(E1, E1R); row3(E0, e0);
            TF gt[2];
#pragma unroll
            for (int c = 0; c < 2; ++c)
            {
                const TF sk = X10(sx, c), sk1 = S1R[c];
                const TF ft_a = rhoh_f * vfl(of, W1R[c], sa[c], sb[c], sk, sk1, sc[c], sd[c]);
                TF ft_d;
                if (SURFACE && f == ks) ft_d = -rhoh_f * sa_.fluxbot[ij + c];
                else if (SURFACE && f == ke) ft_d = -rhoh_f * sa_.fluxtop[ij + c];
                else
                {
                    const TF evisct = h * (X6(e0, c) + E1R[c]) * tPr_i + svisc;
                    ft_d = rhoh_f * evisct * (sk1 - sk) * dzhi_f;
                }
                gt[c] = ft_d - ft_a;
            }
            if (st)
            {
                TF u4[6], sy[7][2], V0R[2], V0P[2], EM0[2], EP0[2];
                row3(U0, u4); col7(S0, sy); pair(V0, V0R); pair(V0 + P, V0P); pair(E0 - P, EM0); pair(E0 + P, EP0);
                TF fx[3], dx_[3];
#pragma unroll
                for (int m = 0; m < 3; ++m)
                {
                    fx[m] = hflux(X4(u4, m), X10(sx, m - 3), X10(sx, m - 2), X10(sx, m - 1), X10(sx, m), X10(sx, m + 1), X10(sx, m + 2));
                    const TF eviscx = h * (X6(e0, m - 1) + X6(e0, m)) * tPr_i + svisc;
                    dx_[m] = eviscx * (X10(sx, m) - X10(sx, m - 1));
                }
#pragma unroll
                for (int c = 0; c < 2; ++c)
                {
                    const TF fn = hflux(V0P[c], sy[1][c], sy[2][c], sy[3][c], sy[4][c], sy[5][c], sy[6][c]);
                    const TF fs = hflux(V0R[c], sy[0][c], sy[1][c], sy[2][c], sy[3][c], sy[4][c], sy[5][c]);
                    const TF eviscn = h * (X6(e0, c) + EP0[c]) * tPr_i + svisc;
                    const TF eviscs = h * (EM0[c] + X6(e0, c)) * tPr_i + svisc;
                    const TF d = (dx_[c + 1] - dx_[c]) * sa_.dxidxi
                               + (eviscn * (sy[4][c] - sy[3][c]) - eviscs * (sy[3][c] - sy[2][c])) * sa_.dyidyi;
                    const TF ts = -(fx[c + 1] - fx[c]) * dxi - (fn - fs) * dyi + d + (gt[c] - gs[c]) * rdzi_k;
                    if (c == 0) old0 += ts; else old1 += ts;
                }
                gstore2(sa_.st + o_k, old0, old1);
            }
            release(k, s0);
            gs[0] = gt[0]; gs[1] = gt[1];
            sa[0] = sb[0]; sa[1] = sb[1]; sb[0] = X10(sx, 0); sb[1] = X10(sx, 1);
            sc[0] = sd[0]; sc[1] = sd[1]; sd[0] = sn0; sd[1] = sn1;
        }
    }
    else
    {
        // ------------------------------------------------------------------ w at face f = k+1 (+ buoyancy)
        TF wa[2], wb[2], we[2], wc[2], wd[2], gw[2] = {0, 0}, thc[2] = {TF(0), TF(0)};
        if (BUOY && !NSC) { thc[0] = colload(a.th, k0, 0); thc[1] = colload(a.th, k0, 1); }
#pragma unroll
        for (int c = 0; c < 2; ++c)
        {
            wa[c] = colload(a.w, k0 - 1, c); wb[c] = colload(a.w, k0, c); we[c] = colload(a.w, k0 + 2, c);
            wc[c] = colload(a.w, k0 + 3, c); wd[c] = colload(a.w, k0 + 4, c);
        }
        for (int k = k0; k < kc1; ++k)
        {
            const int f = k + 1, pl = k - k0;
            const bool st = (k >= kc0) && active && f < ke;
            const long long o_f = ij + (long long)f * kk;
            TF wn0, wn1; colload2(a.w, k + 5, wn0, wn1);
            TF thn0 = TF(0), thn1 = TF(0);          // th[k+1] straight from global when the scalar is not staged (issued early)
            if (BUOY && !NSC) colload2(a.th, k + 1, thn0, thn1);
            TF old0 = 0, old1 = 0;
            if (st) gload2(a.wt + o_f, old0, old1);
            const int oc = vorder(f, ks - 1, ke);
            const TF rho_c = p_rho[pl + 1], dzi_c = p_dzi[pl + 1], rdzhi_f = p_rdzhi[pl + 1], dzhi_f = p_dzhi[pl + 1];
            int s0, s1;
            acquire(k, s0, s1);
            const TF* __restrict__ U0 = plane(0, s0); const TF* __restrict__ U1 = plane(0, s1);
            const TF* __restrict__ V0 = plane(1, s0); const TF* __restrict__ V1 = plane(1, s1);
            const TF* __restrict__ W1 = plane(2, s1);
            const TF* __restrict__ E0 = plane(3, s0); const TF* __restrict__ E1 = plane(3, s1);
            TF wx[12], E1R[2], thk[2] = {TF(0), TF(0)}, th1[2] = {TF(0), TF(0)};
            row10(W1, wx); pair(E1, E1R);
            if (BUOY && NSC) { pair(plane(4, s0), thk); pair(plane(4, s1), th1); }       // th rides in the scalar planes
            else if (BUOY) { thk[0] = thc[0]; thk[1] = thc[1]; th1[0] = thn0; th1[1] = thn1; thc[0] = thn0; thc[1] = thn1; }
            TF gt[2];
#pragma unroll
            for (int c = 0; c < 2; ++c)
            {
                const TF wf = X10(wx, c);
                const TF ft_a = rho_c * vfl(oc, interp2(wf, we[c]), wa[c], wb[c], wf, we[c], wc[c], wd[c]);
                const TF ft_d = rho_c * (E1R[c] + visc) * (we[c] - wf) * dzi_c;
                gt[c] = TF(2.) * ft_d - ft_a;
            }
            if (st)
            {
                TF u4[6], u14[6], e0[6], e16[6], s0r[3], s1r[3];
                row3(U0, u4); row3(U1, u14); row3(E0, e0); row3(E1, e16);
                psum(e0, s0r); psum(e16, s1r);
                TF fx[3], dx_[3];
#pragma unroll
                for (int m = 0; m < 3; ++m)
                {
                    fx[m] = hflux(interp2(X4(u4, m), X4(u14, m)), X10(wx, m - 3), X10(wx, m - 2), X10(wx, m - 1), X10(wx, m), X10(wx, m + 1), X10(wx, m + 2));
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
                    const TF fn = hflux(interp2(V0P[c], V1P[c]), wy[1][c], wy[2][c], wy[3][c], wy[4][c], wy[5][c], wy[6][c]);
                    const TF fs = hflux(interp2(V0R[c], V1R[c]), wy[0][c], wy[1][c], wy[2][c], wy[3][c], wy[4][c], wy[5][c]);
                    const TF eviscn = q * ((X6(e0, c) + X6(e16, c)) + (EP0[c] + EP1[c])) + visc;
                    const TF eviscs = q * ((EM0[c] + X6(e0, c)) + (EM1[c] + X6(e16, c))) + visc;
                    TF tw = -(fx[c + 1] - fx[c]) * dxi - (fn - fs) * dyi
                          + (dx_[c + 1] - dx_[c]) * dxi
                          + (eviscn * ((wy[4][c] - wy[3][c]) * dyi + (V1P[c] - V0P[c]) * dzhi_f)
                           - eviscs * ((wy[3][c] - wy[2][c]) * dyi + (V1R[c] - V0R[c]) * dzhi_f)) * dyi
                          + (gt[c] - gw[c]) * rdzhi_f;
                    if (BUOY) tw += p_gth[pl + 1] * (interp2(thk[c], th1[c]) - p_thh[pl + 1]);
                    if (c == 0) old0 += tw; else old1 += tw;
                }
                gstore2(a.wt + o_f, old0, old1);
            }
            release(k, s0);
            gw[0] = gt[0]; gw[1] = gt[1];
            wa[0] = wb[0]; wa[1] = wb[1]; wb[0] = X10(wx, 0); wb[1] = X10(wx, 1);
            we[0] = wc[0]; we[1] = wc[1]; wc[0] = wd[0]; wc[1] = wd[1]; wd[0] = wn0; wd[1] = wn1;
        }
    }
#undef X10
#undef X6
#undef X4
}

} // namespace mhh
