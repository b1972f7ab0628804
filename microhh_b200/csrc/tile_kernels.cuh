// mhhb200 -- 2.5-D blocked, z-marching tile kernels: the fast path of mhh_dycore_substep.
//
// One CTA owns an xy tile (TX x TY columns, one thread per column) and marches up a chunk of
// levels.  Horizontal neighbours come from shared-memory planes that are streamed in with
// cp.async (LDGSTS) three levels deep, the own column lives in a sliding register window, and
// every vertical face flux is computed once and carried to the next level.  Each field is read
// from HBM once per kernel; halo re-reads are served by L2.
//
// Arithmetic: identical formulas to stencil_kernels.cuh (flux form of reference
// src/advec_2i5.cxx:151-728 and include/diff_kernels.h:144-484, src/thermo_dry.cxx:165-179).
#pragma once
#include "common.cuh"
#include "stencil_kernels.cuh"

namespace mhh {

constexpr int TILE_X = 32;
constexpr int TILE_H = 3;                         // halo of the staged planes (advection stencils)
constexpr int TILE_PX = TILE_X + 2 * TILE_H;      // 38
constexpr int RING = 3;
// TY (tile height = warps per CTA) is a template parameter: 16 -> one 512-thread CTA per SM,
// 8 -> two 256-thread CTAs per SM (fp64 kernels use ~128 registers per thread).
constexpr int tile_plane(int ty, int h = TILE_H) { return (TILE_X + 2 * h) * (ty + 2 * h); }

template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
    else if (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// Stages horizontal planes (tile + halo H) into shared memory with cp.async.  The per-thread
// source/destination offsets are computed once; staging a plane is then NIT predicated copies.
// VEC = elements per cp.async; the host dispatch guarantees that a vector never straddles the
// end of a row (icells and the tile origin are multiples of VEC).
template <typename TF, int VEC, int TY, int H>
struct Stager
{
    static constexpr int PX = TILE_X + 2 * H, PY = TY + 2 * H;
    static constexpr int NV = PX / VEC;
    static constexpr int THREADS = TILE_X * TY;
    static constexpr int NIT = (NV * PY + THREADS - 1) / THREADS;
    int soff[NIT];
    int goff[NIT];      // offset inside a horizontal plane, or -1

    __device__ __forceinline__ void init(const int gi0, const int gj0, const GridDev<TF>& g)
    {
#pragma unroll
        for (int it = 0; it < NIT; ++it)
        {
            const int t = threadIdx.x + it * THREADS;
            const int sy = t / NV;
            const int sx = (t - sy * NV) * VEC;
            const int gj = gj0 + sy, gi = gi0 + sx;
            const bool ok = (t < NV * PY) && (gj < g.jcells) && (gi + VEC <= g.icells);
            soff[it] = sy * PX + sx;
            goff[it] = ok ? gj * g.icells + gi : -1;
        }
    }
    // src_plane = field pointer at the level to stage
    __device__ __forceinline__ void stage(TF* __restrict__ dst, const TF* __restrict__ src_plane) const
    {
#pragma unroll
        for (int it = 0; it < NIT; ++it)
            if (goff[it] >= 0)
                cp_async<VEC * (int)sizeof(TF)>(dst + soff[it], src_plane + goff[it]);
    }
};

template <typename TF>
__device__ __forceinline__ TF vflux_col(const int order, const TF vel, const TF c0, const TF c1, const TF c2,
        const TF c3, const TF c4, const TF c5)
{
    // c0..c5 = column values at f-3 .. f+2 around face/centre f
    if (order == 6) return flux65(vel, c0, c1, c2, c3, c4, c5);
    if (order == 4) return flux43(vel, c1, c2, c3, c4);
    if (order == 2) return flux2(vel, c2, c3);
    return TF(0);
}

template <typename TF>
struct MomTileArgs
{
    MomArgs<TF> m;
    int kchunk;        // levels per CTA in z
};

template <typename TF, bool SURFACE, bool BUOY, int VEC, int TY>
__global__ void __launch_bounds__(TILE_X * TY, 512 / (TILE_X * TY)) mom_tile_kernel(const MomTileArgs<TF> args, const GridDev<TF> g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TF* sm = reinterpret_cast<TF*>(smem_raw);
    constexpr int TILE_Y = TY, TILE_THREADS = TILE_X * TY, TILE_PLANE = tile_plane(TY);
    // layout: [field 0..3][ring][plane]   fields: 0 u, 1 v, 2 w, 3 evisc
    const MomArgs<TF>& a = args.m;
    const int tx = threadIdx.x % TILE_X, ty = threadIdx.x / TILE_X;
    const int i = g.istart + blockIdx.x * TILE_X + tx;
    const int j = g.jstart + blockIdx.y * TILE_Y + ty;
    const int gi0 = g.istart + blockIdx.x * TILE_X - TILE_H;     // >= 0 because igc >= 3
    const int gj0 = g.jstart + blockIdx.y * TILE_Y - TILE_H;
    const bool active = (i < g.iend) && (j < g.jend);
    const int ks = g.kstart, ke = g.kend;
    const int kc0 = ks + blockIdx.z * args.kchunk;
    const int kc1 = min(ke, kc0 + args.kchunk);
    const long long jj = g.icells, kk = g.ijcells;
    // clamp the column position of inactive threads so that their (unused) loads stay in bounds
    const int ic = min(i, g.iend - 1), jc = min(j, g.jend - 1);
    const long long ij = ic + jc * jj;
    const int sidx = (ty + TILE_H) * TILE_PX + (tx + TILE_H);
    const TF dxi = g.dxi, dyi = g.dyi, visc = a.visc;
    const TF q = TF(0.25);

    const TF* flds[4] = {a.u, a.v, a.w, a.evisc};

    // per-level profiles of this chunk (levels kc0-1 .. kc1+1) in shared memory
    const int k0 = kc0 - 1;
    TF* prof = sm + 4 * RING * TILE_PLANE;
    const int nlev = args.kchunk + 3;
    TF* p_rho = prof; TF* p_rhoh = prof + nlev; TF* p_dzi = prof + 2 * nlev; TF* p_dzhi = prof + 3 * nlev; TF* p_thh = prof + 4 * nlev;
    for (int t = threadIdx.x; t < nlev; t += TILE_THREADS)
    {
        const int lev = min(max(k0 + t, 0), g.kcells - 1);
        p_rho[t] = g.rhoref[lev]; p_rhoh[t] = g.rhorefh[lev]; p_dzi[t] = g.dzi[lev]; p_dzhi[t] = g.dzhi[lev];
        p_thh[t] = BUOY ? g.threfh[lev] : TF(1);
    }

    // prologue: planes k0 = kc0-1 and k0+1 into ring slots 0 and 1
    Stager<TF, VEC, TY, TILE_H> stg;
    stg.init(gi0, gj0, g);
    const TF* pl_src[4];       // per field: pointer to the plane that is staged next
#pragma unroll
    for (int f = 0; f < 4; ++f)
    {
        pl_src[f] = flds[f] + (long long)k0 * kk;
        stg.stage(sm + (f * RING + 0) * TILE_PLANE, pl_src[f]);
        stg.stage(sm + (f * RING + 1) * TILE_PLANE, pl_src[f] + kk);
        pl_src[f] += 2 * kk;
    }
    cp_async_commit();
    int s0 = 0;                // ring slot of plane k

    auto colload = [&](const TF* __restrict__ fld, int lev) -> TF {
        return (lev >= 0 && lev < g.kcells) ? fld[ij + (long long)lev * kk] : TF(0);
    };
    // register windows: uc[n] = u[k-2+n], n = 0..5 (k-2..k+3); wc[n] = w[k-1+n] (k-1..k+4)
    TF uc[6], vc[6], wc[6];
#pragma unroll
    for (int n = 0; n < 6; ++n)
    {
        uc[n] = colload(a.u, k0 - 2 + n);
        vc[n] = colload(a.v, k0 - 2 + n);
        wc[n] = colload(a.w, k0 - 1 + n);
    }
    TF thk = TF(0);
    if (BUOY) thk = colload(a.th, k0);

    // carried vertical fluxes through the bottom face (u, v) / bottom centre (w)
    TF fa_u = 0, fd_u = 0, fa_v = 0, fd_v = 0, fa_w = 0, fd_w = 0;

    for (int k = k0; k < kc1; ++k)
    {
        cp_async_wait<0>();          // planes k and k+1 have landed
        __syncthreads();             // ... for every thread; and everyone is done with plane k-1
        // stream in plane k+2 (into the slot of plane k-1) while level k is being computed
        const int s1 = (s0 == RING - 1) ? 0 : s0 + 1;
        const int s2 = (s1 == RING - 1) ? 0 : s1 + 1;
        if (k + 2 < g.kcells)
        {
#pragma unroll
            for (int f = 0; f < 4; ++f) { stg.stage(sm + (f * RING + s2) * TILE_PLANE, pl_src[f]); pl_src[f] += kk; }
        }
        cp_async_commit();
        // global loads whose latency overlaps the arithmetic below: leading-edge column values for the
        // next step and the tendencies that are read-modify-written at the end of this step
        const TF u_new = colload(a.u, k + 4);
        const TF v_new = colload(a.v, k + 4);
        const TF w_new = colload(a.w, k + 5);
        TF th1 = TF(0);
        if (BUOY) th1 = colload(a.th, k + 1);
        const bool store = (k >= kc0);
        const int f = k + 1;                                    // top face of cell k == w level handled here
        const bool st_uv = store && active;
        const bool st_w = store && active && f < ke;
        const long long o_k = ij + (long long)k * kk, o_f = ij + (long long)f * kk;
        TF ut_old = TF(0), vt_old = TF(0), wt_old = TF(0);
        if (st_uv) { ut_old = a.ut[o_k]; vt_old = a.vt[o_k]; }
        if (st_w) wt_old = a.wt[o_f];

        const TF* __restrict__ U0 = sm + (0 * RING + s0) * TILE_PLANE + sidx;
        const TF* __restrict__ U1 = sm + (0 * RING + s1) * TILE_PLANE + sidx;
        const TF* __restrict__ V0 = sm + (1 * RING + s0) * TILE_PLANE + sidx;
        const TF* __restrict__ V1 = sm + (1 * RING + s1) * TILE_PLANE + sidx;
        const TF* __restrict__ W1 = sm + (2 * RING + s1) * TILE_PLANE + sidx;
        const TF* __restrict__ E0 = sm + (3 * RING + s0) * TILE_PLANE + sidx;
        const TF* __restrict__ E1 = sm + (3 * RING + s1) * TILE_PLANE + sidx;
        constexpr int P = TILE_PX;
        const int pl = k - k0;                                  // profile slot of level k
        const TF rho_k = p_rho[pl], rhoh_f = p_rhoh[pl + 1];
        const TF dzi_k = p_dzi[pl], dzhi_f = p_dzhi[pl + 1];

        // ------------------------------------------------------------------ u at cell k
        {
            const int of = vorder(f, ks, ke);
            const TF ft_a = rhoh_f * vflux_col<TF>(of, interp2(W1[-1], W1[0]), uc[0], uc[1], uc[2], uc[3], uc[4], uc[5]);
            TF ft_d;
            if (SURFACE && f == ks) ft_d = -rhoh_f * a.u_fluxbot[ij];
            else if (SURFACE && f == ke) ft_d = -rhoh_f * a.u_fluxtop[ij];
            else
            {
                const TF evisct = q * (E0[-1] + E0[0] + E1[-1] + E1[0]) + visc;
                ft_d = rhoh_f * evisct * ((uc[3] - uc[2]) * dzhi_f + (W1[0] - W1[-1]) * dxi);
            }
            if (st_uv)
            {
                const TF fe = flux65(interp2(U0[0], U0[1]), U0[-2], U0[-1], U0[0], U0[1], U0[2], U0[3]);
                const TF fw = flux65(interp2(U0[-1], U0[0]), U0[-3], U0[-2], U0[-1], U0[0], U0[1], U0[2]);
                const TF fn = flux65(interp2(V0[P - 1], V0[P]), U0[-2 * P], U0[-P], U0[0], U0[P], U0[2 * P], U0[3 * P]);
                const TF fs = flux65(interp2(V0[-1], V0[0]), U0[-3 * P], U0[-2 * P], U0[-P], U0[0], U0[P], U0[2 * P]);
                const TF evisce = E0[0] + visc;
                const TF eviscw = E0[-1] + visc;
                const TF eviscn = q * (E0[-1] + E0[0] + E0[P - 1] + E0[P]) + visc;
                const TF eviscs = q * (E0[-P - 1] + E0[-P] + E0[-1] + E0[0]) + visc;
                const TF d = (evisce * (U0[1] - U0[0]) * dxi - eviscw * (U0[0] - U0[-1]) * dxi) * TF(2.) * dxi
                           + (eviscn * ((U0[P] - U0[0]) * dyi + (V0[P] - V0[P - 1]) * dxi)
                            - eviscs * ((U0[0] - U0[-P]) * dyi + (V0[0] - V0[-1]) * dxi)) * dyi;
                const TF tu = -(fe - fw) * dxi - (fn - fs) * dyi - (ft_a - fa_u) / rho_k * dzi_k
                            + d + (ft_d - fd_u) / rho_k * dzi_k;
                a.ut[o_k] = ut_old + tu;
            }
            fa_u = ft_a; fd_u = ft_d;
        }
        // ------------------------------------------------------------------ v at cell k
        {
            const int of = vorder(f, ks, ke);
            const TF ft_a = rhoh_f * vflux_col<TF>(of, interp2(W1[-P], W1[0]), vc[0], vc[1], vc[2], vc[3], vc[4], vc[5]);
            TF ft_d;
            if (SURFACE && f == ks) ft_d = -rhoh_f * a.v_fluxbot[ij];
            else if (SURFACE && f == ke) ft_d = -rhoh_f * a.v_fluxtop[ij];
            else
            {
                const TF evisct = q * (E0[-P] + E0[0] + E1[-P] + E1[0]) + visc;
                ft_d = rhoh_f * evisct * ((vc[3] - vc[2]) * dzhi_f + (W1[0] - W1[-P]) * dyi);
            }
            if (st_uv)
            {
                const TF fe = flux65(interp2(U0[1 - P], U0[1]), V0[-2], V0[-1], V0[0], V0[1], V0[2], V0[3]);
                const TF fw = flux65(interp2(U0[-P], U0[0]), V0[-3], V0[-2], V0[-1], V0[0], V0[1], V0[2]);
                const TF fn = flux65(interp2(V0[0], V0[P]), V0[-2 * P], V0[-P], V0[0], V0[P], V0[2 * P], V0[3 * P]);
                const TF fs = flux65(interp2(V0[-P], V0[0]), V0[-3 * P], V0[-2 * P], V0[-P], V0[0], V0[P], V0[2 * P]);
                const TF evisce = q * (E0[-P] + E0[0] + E0[1 - P] + E0[1]) + visc;
                const TF eviscw = q * (E0[-1 - P] + E0[-1] + E0[-P] + E0[0]) + visc;
                const TF eviscn = E0[0] + visc;
                const TF eviscs = E0[-P] + visc;
                const TF d = (evisce * ((V0[1] - V0[0]) * dxi + (U0[1] - U0[1 - P]) * dyi)
                            - eviscw * ((V0[0] - V0[-1]) * dxi + (U0[0] - U0[-P]) * dyi)) * dxi
                           + (eviscn * (V0[P] - V0[0]) * dyi - eviscs * (V0[0] - V0[-P]) * dyi) * TF(2.) * dyi;
                const TF tv = -(fe - fw) * dxi - (fn - fs) * dyi - (ft_a - fa_v) / rho_k * dzi_k
                            + d + (ft_d - fd_v) / rho_k * dzi_k;
                a.vt[o_k] = vt_old + tv;
            }
            fa_v = ft_a; fd_v = ft_d;
        }
        // ------------------------------------------------------------------ w at face f = k+1
        {
            // top centre of face f is cell f; order from the distance of the centre to the walls
            const int oc = vorder(f, ks - 1, ke);
            const TF rho_c = p_rho[pl + 1];
            const TF ft_a = rho_c * vflux_col<TF>(oc, interp2(wc[2], wc[3]), wc[0], wc[1], wc[2], wc[3], wc[4], wc[5]);
            const TF ft_d = rho_c * (E1[0] + visc) * (wc[3] - wc[2]) * p_dzi[pl + 1];
            if (st_w)
            {
                const TF fe = flux65(interp2(U0[1], U1[1]), W1[-2], W1[-1], W1[0], W1[1], W1[2], W1[3]);
                const TF fw = flux65(interp2(U0[0], U1[0]), W1[-3], W1[-2], W1[-1], W1[0], W1[1], W1[2]);
                const TF fn = flux65(interp2(V0[P], V1[P]), W1[-2 * P], W1[-P], W1[0], W1[P], W1[2 * P], W1[3 * P]);
                const TF fs = flux65(interp2(V0[0], V1[0]), W1[-3 * P], W1[-2 * P], W1[-P], W1[0], W1[P], W1[2 * P]);
                const TF evisce = q * (E0[0] + E1[0] + E0[1] + E1[1]) + visc;
                const TF eviscw = q * (E0[-1] + E1[-1] + E0[0] + E1[0]) + visc;
                const TF eviscn = q * (E0[0] + E1[0] + E0[P] + E1[P]) + visc;
                const TF eviscs = q * (E0[-P] + E1[-P] + E0[0] + E1[0]) + visc;
                TF tw = -(fe - fw) * dxi - (fn - fs) * dyi - (ft_a - fa_w) / rhoh_f * dzhi_f
                      + (evisce * ((W1[1] - W1[0]) * dxi + (U1[1] - U0[1]) * dzhi_f)
                       - eviscw * ((W1[0] - W1[-1]) * dxi + (U1[0] - U0[0]) * dzhi_f)) * dxi
                      + (eviscn * ((W1[P] - W1[0]) * dyi + (V1[P] - V0[P]) * dzhi_f)
                       - eviscs * ((W1[0] - W1[-P]) * dyi + (V1[0] - V0[0]) * dzhi_f)) * dyi
                      + (ft_d - fd_w) / rhoh_f * TF(2.) * dzhi_f;
                if (BUOY) tw += TF(GRAV) / p_thh[pl + 1] * (interp2(thk, th1) - p_thh[pl + 1]);
                a.wt[o_f] = wt_old + tw;
            }
            fa_w = ft_a; fd_w = ft_d;
        }

        // slide the register windows
#pragma unroll
        for (int n = 0; n < 5; ++n) { uc[n] = uc[n + 1]; vc[n] = vc[n + 1]; wc[n] = wc[n + 1]; }
        uc[5] = u_new; vc[5] = v_new; wc[5] = w_new;
        thk = th1;
        s0 = s1;
    }
    cp_async_wait<0>();
}

inline size_t mom_tile_smem(size_t elem, int kchunk, int ty) { return ((size_t)4 * RING * tile_plane(ty) + (size_t)5 * (kchunk + 3)) * elem; }


// ------------------------------------------------------------------------------------------
// Scalar tendency (Advec_2i5 advec_s + Diff_kernels::diff_c), z-marching.  Only the scalar and
// evisc need horizontal neighbours, so only those two are staged (ring of 2 planes); u, v, w
// enter with coalesced direct loads issued a step ahead.
// ------------------------------------------------------------------------------------------
template <typename TF>
struct ScalTileArgs
{
    ScalArgs<TF> s;
    int kchunk;
};

template <typename TF, bool SURFACE, int VEC, int TY>
__global__ void __launch_bounds__(TILE_X * TY, 512 / (TILE_X * TY)) scal_tile_kernel(const ScalTileArgs<TF> args, const GridDev<TF> g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TF* sm = reinterpret_cast<TF*>(smem_raw);
    constexpr int TILE_Y = TY, TILE_THREADS = TILE_X * TY, TILE_PLANE = tile_plane(TY);
    constexpr int R2 = 2;

    const ScalArgs<TF>& a = args.s;
    const int tx = threadIdx.x % TILE_X, ty = threadIdx.x / TILE_X;
    const int i = g.istart + blockIdx.x * TILE_X + tx;
    const int j = g.jstart + blockIdx.y * TILE_Y + ty;
    const int gi0 = g.istart + blockIdx.x * TILE_X - TILE_H;
    const int gj0 = g.jstart + blockIdx.y * TILE_Y - TILE_H;
    const bool active = (i < g.iend) && (j < g.jend);
    const int ks = g.kstart, ke = g.kend;
    const int kc0 = ks + blockIdx.z * args.kchunk;
    const int kc1 = min(ke, kc0 + args.kchunk);
    const long long jj = g.icells, kk = g.ijcells;
    const int ic = min(i, g.iend - 1), jc = min(j, g.jend - 1);
    const long long ij = ic + jc * jj;
    const int sidx = (ty + TILE_H) * TILE_PX + (tx + TILE_H);
    const TF dxi = g.dxi, dyi = g.dyi, visc = a.visc;
    const TF h = TF(0.5);
    const TF tPr_i = TF(1) / a.tPr;

    const int k0 = kc0 - 1;
    TF* prof = sm + 2 * R2 * TILE_PLANE;
    const int nlev = args.kchunk + 3;
    TF* p_rho = prof; TF* p_rhoh = prof + nlev; TF* p_dzi = prof + 2 * nlev; TF* p_dzhi = prof + 3 * nlev;
    for (int t = threadIdx.x; t < nlev; t += TILE_THREADS)
    {
        const int lev = min(max(k0 + t, 0), g.kcells - 1);
        p_rho[t] = g.rhoref[lev]; p_rhoh[t] = g.rhorefh[lev]; p_dzi[t] = g.dzi[lev]; p_dzhi[t] = g.dzhi[lev];
    }

    Stager<TF, VEC, TY, TILE_H> stg;
    stg.init(gi0, gj0, g);
    const TF* src_s = a.s + (long long)k0 * kk;
    const TF* src_e = a.evisc + (long long)k0 * kk;
    stg.stage(sm + (0 * R2 + 0) * TILE_PLANE, src_s); src_s += kk;
    stg.stage(sm + (1 * R2 + 0) * TILE_PLANE, src_e); src_e += kk;
    cp_async_commit();
    int s0 = 0;

    auto colload = [&](const TF* __restrict__ fld, int lev) -> TF {
        return (lev >= 0 && lev < g.kcells) ? fld[ij + (long long)lev * kk] : TF(0);
    };
    TF sc[6];
#pragma unroll
    for (int n = 0; n < 6; ++n) sc[n] = colload(a.s, k0 - 2 + n);
    TF fa = 0, fd = 0;

    for (int k = k0; k < kc1; ++k)
    {
        cp_async_wait<0>();
        __syncthreads();
        const int s1 = s0 ^ 1;
        if (k + 1 < g.kcells)
        {
            stg.stage(sm + (0 * R2 + s1) * TILE_PLANE, src_s); src_s += kk;
            stg.stage(sm + (1 * R2 + s1) * TILE_PLANE, src_e); src_e += kk;
        }
        cp_async_commit();

        const bool store = (k >= kc0) && active;
        const int f = k + 1;
        const long long o_k = ij + (long long)k * kk;
        const TF s_new = colload(a.s, k + 4);
        const TF e1 = colload(a.evisc, f);
        const TF w_f = colload(a.w, f);
        TF u0 = 0, u1 = 0, v0 = 0, v1 = 0, st_old = 0;
        if (store)
        {
            u0 = a.u[o_k]; u1 = a.u[o_k + 1]; v0 = a.v[o_k]; v1 = a.v[o_k + jj];
            st_old = a.st[o_k];
        }
        const TF* __restrict__ S0 = sm + (0 * R2 + s0) * TILE_PLANE + sidx;
        const TF* __restrict__ E0 = sm + (1 * R2 + s0) * TILE_PLANE + sidx;
        constexpr int P = TILE_PX;
        const int pl = k - k0;
        const TF rho_k = p_rho[pl], rhoh_f = p_rhoh[pl + 1], dzi_k = p_dzi[pl], dzhi_f = p_dzhi[pl + 1];

        const TF ft_a = rhoh_f * vflux_col<TF>(vorder(f, ks, ke), w_f, sc[0], sc[1], sc[2], sc[3], sc[4], sc[5]);
        TF ft_d;
        if (SURFACE && f == ks) ft_d = -rhoh_f * a.fluxbot[ij];
        else if (SURFACE && f == ke) ft_d = -rhoh_f * a.fluxtop[ij];
        else
        {
            const TF evisct = h * (E0[0] + e1) * tPr_i + visc;
            ft_d = rhoh_f * evisct * (sc[3] - sc[2]) * dzhi_f;
        }
        if (store)
        {
            const TF fe = flux65(u1, S0[-2], S0[-1], S0[0], S0[1], S0[2], S0[3]);
            const TF fw = flux65(u0, S0[-3], S0[-2], S0[-1], S0[0], S0[1], S0[2]);
            const TF fn = flux65(v1, S0[-2 * P], S0[-P], S0[0], S0[P], S0[2 * P], S0[3 * P]);
            const TF fs = flux65(v0, S0[-3 * P], S0[-2 * P], S0[-P], S0[0], S0[P], S0[2 * P]);
            const TF evisce = h * (E0[0] + E0[1]) * tPr_i + visc;
            const TF eviscw = h * (E0[-1] + E0[0]) * tPr_i + visc;
            const TF eviscn = h * (E0[0] + E0[P]) * tPr_i + visc;
            const TF eviscs = h * (E0[-P] + E0[0]) * tPr_i + visc;
            const TF d = (evisce * (S0[1] - S0[0]) - eviscw * (S0[0] - S0[-1])) * a.dxidxi
                       + (eviscn * (S0[P] - S0[0]) - eviscs * (S0[0] - S0[-P])) * a.dyidyi;
            const TF ts = -(fe - fw) * dxi - (fn - fs) * dyi - (ft_a - fa) / rho_k * dzi_k
                        + d + (ft_d - fd) / rho_k * dzi_k;
            a.st[o_k] = st_old + ts;
        }
        fa = ft_a; fd = ft_d;
#pragma unroll
        for (int n = 0; n < 5; ++n) sc[n] = sc[n + 1];
        sc[5] = s_new;
        s0 = s1;
    }
    cp_async_wait<0>();
}

inline size_t scal_tile_smem(size_t elem, int kchunk, int ty) { return ((size_t)2 * 2 * tile_plane(ty) + (size_t)4 * (kchunk + 3)) * elem; }


// ------------------------------------------------------------------------------------------
// Eddy viscosity, z-marching: strain^2 + N2 + Smagorinsky-Lilly in one pass (u,v,w,th -> evisc).
// Planes of u, v (levels k, k+1) and w (level k+1) with a one-cell halo; the vertical-shear
// terms of the top face are carried to the next level.
// ------------------------------------------------------------------------------------------
constexpr int EH = 1;
constexpr int ERING = 4;                  // planes in flight per field: levels k, k+1 in use, k+2 and k+3 on their way
constexpr int EPX = TILE_X + 2 * EH;      // 34

template <typename TF>
struct EviscTileArgs
{
    EviscArgs<TF> e;
    const TF* mlen0;
    int kchunk;
};

template <typename TF, bool SURFACE, int VEC, int TY, int MB = 512 / (TILE_X * TY)>
__global__ void __launch_bounds__(TILE_X * TY, MB) evisc_tile_kernel(const EviscTileArgs<TF> args, const GridDev<TF> g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TF* sm = reinterpret_cast<TF*>(smem_raw);
    constexpr int TILE_Y = TY, TILE_THREADS = TILE_X * TY, EPLANE = tile_plane(TY, EH);
    // u: slots 0..3, v: 4..7, w: 8..11  (ring of ERING each)

    const EviscArgs<TF>& a = args.e;
    const int tx = threadIdx.x % TILE_X, ty = threadIdx.x / TILE_X;
    const int i = g.istart + blockIdx.x * TILE_X + tx;
    const int j = g.jstart + blockIdx.y * TILE_Y + ty;
    const int gi0 = g.istart + blockIdx.x * TILE_X - EH;
    const int gj0 = g.jstart + blockIdx.y * TILE_Y - EH;
    const bool active = (i < g.iend) && (j < g.jend);
    const int ks = g.kstart, ke = g.kend;
    const int kc0 = ks + blockIdx.z * args.kchunk;
    const int kc1 = min(ke, kc0 + args.kchunk);
    const long long jj = g.icells, kk = g.ijcells;
    const int ic = min(i, g.iend - 1), jc = min(j, g.jend - 1);
    const long long ij = ic + jc * jj;
    const int sidx = (ty + EH) * EPX + (tx + EH);
    const TF dxi = g.dxi, dyi = g.dyi;
    const TF e8 = TF(0.125);

    const int k0 = kc0 - 1;
    constexpr int RING = ERING;
    TF* prof = sm + 3 * RING * EPLANE;
    const int nlev = args.kchunk + 3;
    TF* p_dzi = prof; TF* p_dzhi = prof + nlev; TF* p_m0 = prof + 2 * nlev; TF* p_z = prof + 3 * nlev; TF* p_gth = prof + 4 * nlev;
    for (int t = threadIdx.x; t < nlev; t += TILE_THREADS)
    {
        const int lev = min(max(k0 + t, 0), g.kcells - 1);
        p_dzi[t] = g.dzi[lev]; p_dzhi[t] = g.dzhi[lev]; p_z[t] = g.z[lev];
        const TF m0 = a.cs * args.mlen0[lev];
        p_m0[t] = m0 * m0;
        p_gth[t] = (a.n2mode == 1) ? TF(GRAV) / g.thref[lev] : TF(0);
    }
    const TF* flds[3] = {a.u, a.v, a.w};
    Stager<TF, VEC, TY, EH> stg;
    stg.init(gi0, gj0, g);
    const TF* pl_src[3];
#pragma unroll
    for (int f = 0; f < 3; ++f)
    {
        pl_src[f] = flds[f] + (long long)k0 * kk;
        stg.stage(sm + (f * RING + 0) * EPLANE, pl_src[f]);
        stg.stage(sm + (f * RING + 1) * EPLANE, pl_src[f] + kk);
        pl_src[f] += 2 * kk;
    }
    cp_async_commit();
    if (k0 + 2 < g.kcells)
    {
#pragma unroll
        for (int f = 0; f < 3; ++f) { stg.stage(sm + (f * RING + 2) * EPLANE, pl_src[f]); pl_src[f] += kk; }
    }
    cp_async_commit();
    int s0 = 0;

    auto colload = [&](const TF* __restrict__ fld, int lev) -> TF {
        return (lev >= 0 && lev < g.kcells) ? fld[ij + (long long)lev * kk] : TF(0);
    };
    // th column window: level k+2 is requested one iteration before it is needed, so the load latency is off the
    // critical path (ncu: 29 % of the stall samples sat on the consumer of the same-iteration load)
    TF th_m = 0, th_c = 0, th_p = 0;
    if (a.n2mode == 1) { th_m = colload(a.th, k0 - 1); th_c = colload(a.th, k0); th_p = colload(a.th, k0 + 1); }
    TF z0 = TF(0), dudz_b = 0, dvdz_b = 0, dbdz_b = 0;
    if (SURFACE) { z0 = a.z0m[ij]; if (kc0 == ks) { dudz_b = a.dudz[ij]; dvdz_b = a.dvdz[ij]; dbdz_b = a.dbdz[ij]; } }

    // carried top-face shear terms: T at (i, f), (i+1, f); R at (j, f), (j+1, f); plus their w-only parts
    TF t0 = 0, t1 = 0, r0 = 0, r1 = 0, tw0 = 0, tw1 = 0, rw0 = 0, rw1 = 0;

    for (int k = k0; k < kc1; ++k)
    {
        cp_async_wait<1>();          // everything but the newest group (level k+2) has landed: planes k and k+1 are ready
        __syncthreads();
        const int s1 = (s0 + 1) & (RING - 1);
        const int s3 = (s0 + 3) & (RING - 1);          // held level k-1, which every thread finished with before the barrier
        if (k + 3 < g.kcells)
        {
#pragma unroll
            for (int f = 0; f < 3; ++f) { stg.stage(sm + (f * RING + s3) * EPLANE, pl_src[f]); pl_src[f] += kk; }
        }
        cp_async_commit();
        const bool store = (k >= kc0) && active;
        const long long o_k = ij + (long long)k * kk;
        TF th_n = 0, n2v = 0;
        if (a.n2mode == 1) th_n = colload(a.th, k + 2);
        else if (store) n2v = a.n2[o_k];

        const TF* __restrict__ U0 = sm + (0 * RING + s0) * EPLANE + sidx;
        const TF* __restrict__ U1 = sm + (0 * RING + s1) * EPLANE + sidx;
        const TF* __restrict__ V0 = sm + (1 * RING + s0) * EPLANE + sidx;
        const TF* __restrict__ V1 = sm + (1 * RING + s1) * EPLANE + sidx;
        const TF* __restrict__ W0 = sm + (2 * RING + s0) * EPLANE + sidx;
        const TF* __restrict__ W1 = sm + (2 * RING + s1) * EPLANE + sidx;
        constexpr int P = EPX;
        const int pl = k - k0;
        const TF dzhi_f = p_dzhi[pl + 1];

        // top-face terms (face f = k+1)
        const TF wx0 = (W1[0] - W1[-1]) * dxi, wx1 = (W1[1] - W1[0]) * dxi;
        const TF wy0 = (W1[0] - W1[-P]) * dyi, wy1 = (W1[P] - W1[0]) * dyi;
        const TF nt0 = (U1[0] - U0[0]) * dzhi_f + wx0;
        const TF nt1 = (U1[1] - U0[1]) * dzhi_f + wx1;
        const TF nr0 = (V1[0] - V0[0]) * dzhi_f + wy0;
        const TF nr1 = (V1[P] - V0[P]) * dzhi_f + wy1;

        if (store)
        {
            TF s = pow2((U0[1] - U0[0]) * dxi) + pow2((V0[P] - V0[0]) * dyi) + pow2((W1[0] - W0[0]) * p_dzi[pl]);
            s += e8 * pow2((U0[0] - U0[-P]) * dyi + (V0[0] - V0[-1]) * dxi);
            s += e8 * pow2((U0[1] - U0[1 - P]) * dyi + (V0[1] - V0[0]) * dxi);
            s += e8 * pow2((U0[P] - U0[0]) * dyi + (V0[P] - V0[P - 1]) * dxi);
            s += e8 * pow2((U0[1 + P] - U0[1]) * dyi + (V0[1 + P] - V0[P]) * dxi);
            const bool bottom_mo = SURFACE && (k == ks);
            if (bottom_mo)
            {
                s += TF(0.5) * pow2(dudz_b);
                s += e8 * pow2(tw0); s += e8 * pow2(tw1); s += e8 * pow2(wx0); s += e8 * pow2(wx1);
                s += TF(0.5) * pow2(dvdz_b);
                s += e8 * pow2(rw0); s += e8 * pow2(rw1); s += e8 * pow2(wy0); s += e8 * pow2(wy1);
            }
            else
            {
                s += e8 * pow2(t0); s += e8 * pow2(t1); s += e8 * pow2(nt0); s += e8 * pow2(nt1);
                s += e8 * pow2(r0); s += e8 * pow2(r1); s += e8 * pow2(nr0); s += e8 * pow2(nr1);
            }
            const TF s2 = (TF)((double)(TF(2.) * s) + DSMALL);
            TF n2;
            if (bottom_mo) n2 = dbdz_b;
            else if (a.n2mode == 1) n2 = p_gth[pl] * TF(0.5) * (th_p - th_m) * p_dzi[pl];
            else n2 = n2v;
            TF rit = n2 / (s2 * a.tPr);
            rit = rit < TF(1. - DSMALL) ? rit : TF(1. - DSMALL);
            TF m2 = p_m0[pl];
            if (SURFACE && a.mason)
            {
                const TF t = TF(KAPPA) * (p_z[pl] + z0);
                const TF t2 = t * t;
                m2 = m2 * t2 / (m2 + t2);       // == 1/(1/mlen0^2 + 1/(kappa (z+z0))^2)
            }
            a.evisc[o_k] = m2 * sqrtf_(s2 * (TF(1.) - rit));
        }
        t0 = nt0; t1 = nt1; r0 = nr0; r1 = nr1; tw0 = wx0; tw1 = wx1; rw0 = wy0; rw1 = wy1;
        th_m = th_c; th_c = th_p; th_p = th_n;
        s0 = s1;
    }
    cp_async_wait<0>();
}

inline size_t evisc_tile_smem(size_t elem, int kchunk, int ty) { return ((size_t)3 * ERING * tile_plane(ty, EH) + (size_t)5 * (kchunk + 3)) * elem; }

} // namespace mhh
