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
constexpr int TILE_Y = 16;
constexpr int TILE_H = 3;                         // halo of the staged planes
constexpr int TILE_PX = TILE_X + 2 * TILE_H;      // 38
constexpr int TILE_PY = TILE_Y + 2 * TILE_H;      // 22
constexpr int TILE_PLANE = TILE_PX * TILE_PY;     // 836 elements
constexpr int TILE_THREADS = TILE_X * TILE_Y;     // 512
constexpr int RING = 3;

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

// Stage one horizontal plane (tile + halo) of `fld` at level `lev` into smem.
// VEC = elements per cp.async (alignment guaranteed by the caller's dispatch).
template <typename TF, int VEC>
__device__ __forceinline__ void stage_plane(TF* __restrict__ dst, const TF* __restrict__ fld, const int lev,
        const int gi0, const int gj0, const GridDev<TF>& g)
{
    if (lev < 0 || lev >= g.kcells) return;       // never consumed
    const TF* src = fld + (long long)lev * g.ijcells;
    constexpr int NV = TILE_PX / VEC;             // vectors per row
    for (int t = threadIdx.x; t < NV * TILE_PY; t += TILE_THREADS)
    {
        const int sy = t / NV;
        const int sx = (t - sy * NV) * VEC;
        const int gj = gj0 + sy;
        const int gi = gi0 + sx;
        if (gj < g.jcells && gi + VEC <= g.icells)
            cp_async<VEC * (int)sizeof(TF)>(dst + sy * TILE_PX + sx, src + (long long)gj * g.icells + gi);
        else if (gj < g.jcells)
        {
            for (int v = 0; v < VEC; ++v)
                if (gi + v < g.icells) dst[sy * TILE_PX + sx + v] = src[(long long)gj * g.icells + gi + v];
        }
    }
}

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

template <typename TF, bool SURFACE, bool BUOY, int VEC>
__global__ void __launch_bounds__(TILE_THREADS, 1) mom_tile_kernel(const MomTileArgs<TF> args, const GridDev<TF> g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TF* sm = reinterpret_cast<TF*>(smem_raw);
    // layout: [field 0..3][ring][plane]   fields: 0 u, 1 v, 2 w, 3 evisc
    auto plane = [&](int fld, int lev) -> TF* { return sm + ((fld * RING) + ((lev + RING) % RING)) * TILE_PLANE; };

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

    // prologue: planes k0 = kc0-1 and k0+1
    const int k0 = kc0 - 1;
#pragma unroll
    for (int f = 0; f < 4; ++f) stage_plane<TF, VEC>(plane(f, k0), flds[f], k0, gi0, gj0, g);
    cp_async_commit();
#pragma unroll
    for (int f = 0; f < 4; ++f) stage_plane<TF, VEC>(plane(f, k0 + 1), flds[f], k0 + 1, gi0, gj0, g);
    cp_async_commit();

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
        // stream in plane k+2 while level k is being computed
#pragma unroll
        for (int f = 0; f < 4; ++f) stage_plane<TF, VEC>(plane(f, k + 2), flds[f], k + 2, gi0, gj0, g);
        cp_async_commit();
        // leading-edge column values for the next step (latency overlaps the arithmetic below)
        const TF u_new = colload(a.u, k + 4);
        const TF v_new = colload(a.v, k + 4);
        const TF w_new = colload(a.w, k + 5);
        TF th1 = TF(0);
        if (BUOY) th1 = colload(a.th, k + 1);

        cp_async_wait<1>();          // planes k and k+1 have landed (only k+2 may be in flight)
        __syncthreads();

        const TF* __restrict__ U0 = plane(0, k) + sidx;
        const TF* __restrict__ U1 = plane(0, k + 1) + sidx;
        const TF* __restrict__ V0 = plane(1, k) + sidx;
        const TF* __restrict__ V1 = plane(1, k + 1) + sidx;
        const TF* __restrict__ W1 = plane(2, k + 1) + sidx;
        const TF* __restrict__ E0 = plane(3, k) + sidx;
        const TF* __restrict__ E1 = plane(3, k + 1) + sidx;
        constexpr int P = TILE_PX;
        const bool store = (k >= kc0);
        const int f = k + 1;                                    // top face of cell k == w level handled here
        const int kr = max(k, 0);                               // profile index guard for the warm-up level
        const TF rho_k = g.rhoref[kr], rhoh_f = g.rhorefh[f];
        const TF dzi_k = g.dzi[kr], dzhi_f = g.dzhi[f];

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
            if (store && active)
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
                a.ut[ij + (long long)k * kk] += tu;
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
            if (store && active)
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
                a.vt[ij + (long long)k * kk] += tv;
            }
            fa_v = ft_a; fd_v = ft_d;
        }
        // ------------------------------------------------------------------ w at face f = k+1
        {
            // top centre of face f is cell f; order from the distance of the centre to the walls
            const int oc = vorder(f, ks - 1, ke);
            const TF rho_c = g.rhoref[f];
            const TF ft_a = rho_c * vflux_col<TF>(oc, interp2(wc[2], wc[3]), wc[0], wc[1], wc[2], wc[3], wc[4], wc[5]);
            const TF ft_d = rho_c * (E1[0] + visc) * (wc[3] - wc[2]) * g.dzi[f];
            if (store && active && f < ke)
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
                if (BUOY) tw += TF(GRAV) / g.threfh[f] * (interp2(thk, th1) - g.threfh[f]);
                a.wt[ij + (long long)f * kk] += tw;
            }
            fa_w = ft_a; fd_w = ft_d;
        }

        // slide the register windows
#pragma unroll
        for (int n = 0; n < 5; ++n) { uc[n] = uc[n + 1]; vc[n] = vc[n + 1]; wc[n] = wc[n + 1]; }
        uc[5] = u_new; vc[5] = v_new; wc[5] = w_new;
        thk = th1;
        __syncthreads();             // everyone is done with plane k before it is overwritten (k+3 -> same slot)
    }
    cp_async_wait<0>();
}

constexpr size_t mom_tile_smem(size_t elem) { return (size_t)4 * RING * TILE_PLANE * elem; }

} // namespace mhh
