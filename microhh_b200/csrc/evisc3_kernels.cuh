// mhhb200 -- TMA-staged, warp-specialised eddy viscosity (Diff_smag2::exec_viscosity: strain^2 + N2 + Smagorinsky-Lilly in
// one z-marching pass, u, v, w, th -> evisc).  Same arithmetic as evisc_tile_kernel (tile_kernels.cuh); what changes is
// who moves the data: a producer warp issues one TMA box per field and level into a ring of shared-memory planes, guarded
// by full / empty mbarriers per ring slot (the cp.async stager of evisc_tile_kernel spent half of its instructions on
// address arithmetic and was held at 64 registers to keep four CTAs per SM resident).  A consumer warp owns one row of the
// 64-wide tile and every lane two points 32 columns apart, so all shared-memory reads are unit-stride 8-byte (conflict free)
// and the per-level overhead (barrier wait, th column, loop) is paid once for two points.
//
// Reference behaviour restated (never copied): calc_strain2 / calc_evisc, src/diff_smag2.cxx:49-310; calc_N2,
// src/thermo_dry.cxx (N2 from the th column).  The halo HL only serves the 16-byte alignment of the TMA box origin
// (istart - HL); the stencil itself needs one cell.
#pragma once
#include "tile3_kernels.cuh"
#include "tile_kernels.cuh"

namespace mhh {

// NPL = points per lane (1 or 2): the tile is 32 * NPL columns wide
constexpr int e3_w(int npl) { return 32 * npl; }
constexpr int e3_px(int hl, int npl) { return e3_w(npl) + 2 * hl; }
constexpr int e3_plane(int ty, int elem, int hl, int npl) { return (e3_px(hl, npl) * (ty + 2) * elem + 127) / 128 * 128 / elem; }
constexpr int e3_box_bytes(int ty, int elem, int hl, int npl) { return e3_px(hl, npl) * (ty + 2) * elem; }
inline size_t evisc3_smem(size_t elem, int kchunk, int ty, int hl, int npl, int ring)
{ return 128 + ((size_t)4 * ring * e3_plane(ty, (int)elem, hl, npl) + (size_t)5 * (kchunk + 3) + (size_t)32 * npl * ty) * elem + 128; }

// halo that puts the box origin istart - hl on a 16-byte boundary (0 = none does: use the cp.async kernel)
inline int e3_pick_hl(int istart, int elem)
{
    if (elem == 8) return (istart % 2 == 0) ? 2 : 1;
    if (istart % 4 == 0) return 4;
    if (istart % 4 == 2) return 2;
    return 0;
}

template <typename TF, bool SURFACE, int TY, int HL, int MB, int NPL, int RING>
__global__ void __launch_bounds__(32 * (TY + 1), MB)
evisc3_kernel(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_v,
              const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_th,
              const EviscTileArgs<TF> args, const GridDev<TF> g)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sbase = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    TF* sm = reinterpret_cast<TF*>(sbase + 128);
    constexpr int P = e3_px(HL, NPL), PLANE = e3_plane(TY, (int)sizeof(TF), HL, NPL), E3_W = e3_w(NPL);
    constexpr unsigned PLANE_BYTES = PLANE * sizeof(TF), BOX_BYTES = e3_box_bytes(TY, (int)sizeof(TF), HL, NPL);
    constexpr int NT = 32 * (TY + 1);

    const EviscArgs<TF>& a = args.e;
    const int warp = threadIdx.x >> 5, tx = threadIdx.x & 31;
    const int gi0 = g.istart + blockIdx.x * E3_W - HL;
    const int gj0 = g.jstart + blockIdx.y * TY - 1;
    const int ks = g.kstart, ke = g.kend;
    const int kc0 = ks + blockIdx.z * args.kchunk;
    const int kc1 = min(ke, kc0 + args.kchunk);
    const int k0 = kc0 - 1;
    const long long jj = g.icells, kk = g.ijcells;

    // the th planes ride along (fourth field) when N2 is derived from th: its column then comes out of shared memory too
    const bool th_tma = a.n2mode == 1;
    TF* prof = sm + 4 * RING * PLANE;
    const int nlev = args.kchunk + 3;
    TF* p_dzi = prof; TF* p_dzhi = prof + nlev; TF* p_m0 = prof + 2 * nlev; TF* p_z = prof + 3 * nlev; TF* p_gth = prof + 4 * nlev;
    TF* p_z0 = prof + 5 * nlev;           // z0m of the tile, one slot per (consumer thread, point): a hand-made register spill
    for (int t = threadIdx.x; t < nlev; t += NT)
    {
        const int lev = min(max(k0 + t, 0), g.kcells - 1);
        p_dzi[t] = g.dzi[lev]; p_dzhi[t] = g.dzhi[lev]; p_z[t] = g.z[lev];
        const TF m0 = a.cs * args.mlen0[lev];
        p_m0[t] = m0 * m0;
        p_gth[t] = (a.n2mode == 1) ? TF(GRAV) / g.thref[lev] : TF(0);
    }
    const unsigned full0 = smem_u32(sbase), empty0 = full0 + 8 * RING;
    const unsigned pl0 = smem_u32(sm);
    if (threadIdx.x == 0)
    {
        for (int s = 0; s < RING; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, TY); }
        mbar_fence_init();
    }
    __syncthreads();

    // ================================================================== producer warp
    if (warp == TY)
    {
        if (tx == 0)
        {
            for (int lev = k0; lev <= kc1; ++lev)
            {
                const int n = lev - k0, slot = n % RING, use = n / RING;
                if (use > 0) mbar_wait(empty0 + 8 * slot, (use - 1) & 1);
                const unsigned bar = full0 + 8 * slot;
                mbar_expect_tx(bar, (th_tma ? 4 : 3) * BOX_BYTES);
                tma_load_3d(pl0 + (0 * RING + slot) * PLANE_BYTES, &tm_u, bar, gi0, gj0, lev);
                tma_load_3d(pl0 + (1 * RING + slot) * PLANE_BYTES, &tm_v, bar, gi0, gj0, lev);
                tma_load_3d(pl0 + (2 * RING + slot) * PLANE_BYTES, &tm_w, bar, gi0, gj0, lev);
                if (th_tma) tma_load_3d(pl0 + (3 * RING + slot) * PLANE_BYTES, &tm_th, bar, gi0, gj0, lev);
            }
        }
        return;
    }

    // ================================================================== consumer warps: row = warp, columns tx and tx + 32
    const int ty = warp;
    const int j = g.jstart + blockIdx.y * TY + ty;
    const int jc = min(j, g.jend - 1);
    bool active[NPL]; int ij[NPL]; int sidx[NPL];            // ij: offset inside one horizontal plane (fits 32 bits)
#pragma unroll
    for (int c = 0; c < NPL; ++c)
    {
        const int i = g.istart + blockIdx.x * E3_W + tx + 32 * c;
        active[c] = (i < g.iend) && (j < g.jend);
        ij[c] = min(i, g.iend - 1) + jc * (int)jj;
        sidx[c] = (ty + 1) * P + HL + tx + 32 * c;
    }
    const TF dxi = g.dxi, dyi = g.dyi;
    const TF e8 = TF(0.125);
    const TF tPr_i = TF(1.) / a.tPr;

    auto colload = [&](const TF* __restrict__ fld, int lev, int c) -> TF {
        return (lev >= 0 && lev < g.kcells) ? fld[ij[c] + (long long)lev * kk] : TF(0);
    };
    // th column window (levels k-1, k) in registers; level k+1 is read from the staged plane k+1
    TF th_m[NPL] = {}, th_c[NPL] = {};
#pragma unroll
    for (int c = 0; c < NPL; ++c)
    {
        if (a.n2mode == 1) { th_m[c] = colload(a.th, k0 - 1, c); th_c[c] = colload(a.th, k0, c); }
        if (SURFACE) p_z0[(c * TY + ty) * 32 + tx] = a.z0m[ij[c]];
    }
    // carried top-face shear terms: T at (i, f), (i+1, f); R at (j, f), (j+1, f).  (Their w-only parts, which the surface
    // row needs for its bottom face, are recomputed there from plane k instead of being carried through every level.)
    TF t0[NPL] = {}, t1[NPL] = {}, r0[NPL] = {}, r1[NPL] = {};

    for (int k = k0; k < kc1; ++k)
    {
        const int n = k - k0, s0 = n % RING, s1 = (n + 1) % RING;
        TF n2v[NPL] = {};
#pragma unroll
        for (int c = 0; c < NPL; ++c)
            if (a.n2mode != 1 && k >= kc0 && active[c]) n2v[c] = a.n2[ij[c] + (long long)k * kk];
        const TF dzhi_f = p_dzhi[n + 1], dzi_k = p_dzi[n];
        if (n == 0) mbar_wait(full0, 0);
        mbar_wait(full0 + 8 * s1, ((n + 1) / RING) & 1);
#pragma unroll
        for (int c = 0; c < NPL; ++c)
        {
            const TF* __restrict__ U0 = sm + (0 * RING + s0) * PLANE + sidx[c];
            const TF* __restrict__ U1 = sm + (0 * RING + s1) * PLANE + sidx[c];
            const TF* __restrict__ V0 = sm + (1 * RING + s0) * PLANE + sidx[c];
            const TF* __restrict__ V1 = sm + (1 * RING + s1) * PLANE + sidx[c];
            const TF* __restrict__ W0 = sm + (2 * RING + s0) * PLANE + sidx[c];
            const TF* __restrict__ W1 = sm + (2 * RING + s1) * PLANE + sidx[c];
            const TF th_p = th_tma ? sm[(3 * RING + s1) * PLANE + sidx[c]] : TF(0);
            // top-face terms (face f = k+1)
            const TF wx0 = (W1[0] - W1[-1]) * dxi, wx1 = (W1[1] - W1[0]) * dxi;
            const TF wy0 = (W1[0] - W1[-P]) * dyi, wy1 = (W1[P] - W1[0]) * dyi;
            const TF nt0 = (U1[0] - U0[0]) * dzhi_f + wx0;
            const TF nt1 = (U1[1] - U0[1]) * dzhi_f + wx1;
            const TF nr0 = (V1[0] - V0[0]) * dzhi_f + wy0;
            const TF nr1 = (V1[P] - V0[P]) * dzhi_f + wy1;
            if (k >= kc0 && active[c])
            {
                const long long o_k = ij[c] + (long long)k * kk;
                TF s = pow2((U0[1] - U0[0]) * dxi) + pow2((V0[P] - V0[0]) * dyi) + pow2((W1[0] - W0[0]) * dzi_k);
                s += e8 * pow2((U0[0] - U0[-P]) * dyi + (V0[0] - V0[-1]) * dxi);
                s += e8 * pow2((U0[1] - U0[1 - P]) * dyi + (V0[1] - V0[0]) * dxi);
                s += e8 * pow2((U0[P] - U0[0]) * dyi + (V0[P] - V0[P - 1]) * dxi);
                s += e8 * pow2((U0[1 + P] - U0[1]) * dyi + (V0[1 + P] - V0[P]) * dxi);
                const bool bottom_mo = SURFACE && (k == ks);
                if (bottom_mo)
                {
                    const TF tw0 = (W0[0] - W0[-1]) * dxi, tw1 = (W0[1] - W0[0]) * dxi;       // bottom face: w of level k
                    const TF rw0 = (W0[0] - W0[-P]) * dyi, rw1 = (W0[P] - W0[0]) * dyi;
                    s += TF(0.5) * pow2(a.dudz[ij[c]]);
                    s += e8 * pow2(tw0); s += e8 * pow2(tw1); s += e8 * pow2(wx0); s += e8 * pow2(wx1);
                    s += TF(0.5) * pow2(a.dvdz[ij[c]]);
                    s += e8 * pow2(rw0); s += e8 * pow2(rw1); s += e8 * pow2(wy0); s += e8 * pow2(wy1);
                }
                else
                {
                    s += e8 * pow2(t0[c]); s += e8 * pow2(t1[c]); s += e8 * pow2(nt0); s += e8 * pow2(nt1);
                    s += e8 * pow2(r0[c]); s += e8 * pow2(r1[c]); s += e8 * pow2(nr0); s += e8 * pow2(nr1);
                }
                const TF s2 = (TF)((double)(TF(2.) * s) + DSMALL);
                TF n2;
                if (bottom_mo) n2 = a.dbdz[ij[c]];
                else if (a.n2mode == 1) n2 = p_gth[n] * TF(0.5) * (th_p - th_m[c]) * dzi_k;
                else n2 = n2v[c];
                // S2 (1 - min(Ri / tPr, 1 - dsmall)) without the division: max(S2 - N2 / tPr, S2 (1 - (1 - dsmall)))
                const TF lo = s2 * (TF(1.) - TF(1. - DSMALL)), sr = s2 - n2 * tPr_i;
                TF m2 = p_m0[n];
                if (SURFACE && a.mason)
                {
                    const TF t = TF(KAPPA) * (p_z[n] + p_z0[(c * TY + ty) * 32 + tx]);
                    const TF t2 = t * t;
                    m2 = m2 * t2 / (m2 + t2);       // == 1/(1/mlen0^2 + 1/(kappa (z+z0))^2)
                }
                a.evisc[o_k] = m2 * sqrtf_(sr > lo ? sr : lo);
            }
            t0[c] = nt0; t1[c] = nt1; r0[c] = nr0; r1[c] = nr1;
            th_m[c] = th_c[c]; th_c[c] = th_p;
        }
        __syncwarp();
        if (tx == 0) mbar_arrive(empty0 + 8 * s0);       // this warp is done with plane k
    }
}

} // namespace mhh
