// mhhb200 -- Advec_2i4 and Advec_2i62: 2nd-order flux divergence of CENTRED higher-order interpolations (no upwind term).
//   Advec_2i4   src/advec_2i4.cxx:53-518    interp4c in x, y, z; on the vertical faces next to a wall interp2, through the wall
//                                           no flux (the reference writes those five row variants out; here the face picks its order)
//   Advec_2i62  src/advec_2i62.cxx:59-306   interp6_ws in x and y, interp2 in z on every level
// Template H = 4 | 6 selects the scheme.  One thread per grid point, u, v, w fused in one kernel and one kernel per scalar, like
// order2_kernels.cuh: each of u, v, w comes from HBM once per launch, the +-3 neighbours are L1/L2 hits.  Algorithmic traffic
// 9 array passes (R u, v, w + RMW ut, vt, wt) for the momentum kernel, 6 per scalar (R s, u, v, w + RMW st).
#pragma once
#include "common.cuh"
#include "order4_kernels.cuh"

namespace mhh {

// q interpolated to the LOWER face of cell o along stride s
template <typename TF, int H>
__device__ __forceinline__ TF i2x_face(const TF* __restrict__ q, const long long o, const long long s)
{
    if (H == 4) return i4m(q[o - 2 * s], q[o - s], q[o], q[o + s]);
    return interp6_ws(q[o - 3 * s], q[o - 2 * s], q[o - s], q[o], q[o + s], q[o + 2 * s]);
}

// a cell-centred field (u, v, scalar) interpolated to the face BELOW level k; `ok` = false: no flux through that face (2i4 at the walls)
template <typename TF, int H>
__device__ __forceinline__ TF i2x_vface(const TF* __restrict__ q, const long long o, const long long kk, const int k, const int ks, const int ke, bool& ok)
{
    ok = true;
    if (H == 6) return interp2(q[o - kk], q[o]);
    if (k <= ks || k >= ke) { ok = false; return TF(0); }
    if (k == ks + 1 || k == ke - 1) return interp2(q[o - kk], q[o]);
    return i4m(q[o - 2 * kk], q[o - kk], q[o], q[o + kk]);
}

// w interpolated to cell centre c (between faces c and c+1); o = index of face c
template <typename TF, int H>
__device__ __forceinline__ TF i2x_wcentre(const TF* __restrict__ w, const long long o, const long long kk, const int c, const int ks, const int ke)
{
    if (H == 6 || c == ks || c == ke - 1) return interp2(w[o], w[o + kk]);
    return i4m(w[o - kk], w[o], w[o + kk], w[o + 2 * kk]);
}

template <typename TF>
struct Adv2iArgs
{
    TF* ut; TF* vt; TF* wt;
    const TF* u; const TF* v; const TF* w;
};

template <typename TF, int H>
__global__ void __launch_bounds__(256) adv2i_uvw_kernel(const Adv2iArgs<TF> a, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ijk = i + j * jj + k * kk;
    const TF* __restrict__ u = a.u; const TF* __restrict__ v = a.v; const TF* __restrict__ w = a.w;
    const TF dxi = g.dxi, dyi = g.dyi;
    const int ks = g.kstart, ke = g.kend;
    const TF rhoh1 = g.rhorefh[k + 1], rhoh0 = g.rhorefh[k], rho = g.rhoref[k], dzi = g.dzi[k];
    bool okt, okb;
    {
        const TF top = i2x_vface<TF, H>(u, ijk + kk, kk, k + 1, ks, ke, okt), bot = i2x_vface<TF, H>(u, ijk, kk, k, ks, ke, okb);
        const TF ft = okt ? rhoh1 * interp2(w[ijk - 1 + kk], w[ijk + kk]) * top : TF(0);
        const TF fb = okb ? rhoh0 * interp2(w[ijk - 1], w[ijk]) * bot : TF(0);
        a.ut[ijk] += - (interp2(u[ijk], u[ijk + 1]) * i2x_face<TF, H>(u, ijk + 1, 1)
                      - interp2(u[ijk - 1], u[ijk]) * i2x_face<TF, H>(u, ijk, 1)) * dxi
                     - (interp2(v[ijk - 1 + jj], v[ijk + jj]) * i2x_face<TF, H>(u, ijk + jj, jj)
                      - interp2(v[ijk - 1], v[ijk]) * i2x_face<TF, H>(u, ijk, jj)) * dyi
                     - (ft - fb) / rho * dzi;
    }
    {
        const TF top = i2x_vface<TF, H>(v, ijk + kk, kk, k + 1, ks, ke, okt), bot = i2x_vface<TF, H>(v, ijk, kk, k, ks, ke, okb);
        const TF ft = okt ? rhoh1 * interp2(w[ijk - jj + kk], w[ijk + kk]) * top : TF(0);
        const TF fb = okb ? rhoh0 * interp2(w[ijk - jj], w[ijk]) * bot : TF(0);
        a.vt[ijk] += - (interp2(u[ijk + 1 - jj], u[ijk + 1]) * i2x_face<TF, H>(v, ijk + 1, 1)
                      - interp2(u[ijk - jj], u[ijk]) * i2x_face<TF, H>(v, ijk, 1)) * dxi
                     - (interp2(v[ijk], v[ijk + jj]) * i2x_face<TF, H>(v, ijk + jj, jj)
                      - interp2(v[ijk - jj], v[ijk]) * i2x_face<TF, H>(v, ijk, jj)) * dyi
                     - (ft - fb) / rho * dzi;
    }
    if (k > ks)          // w tendencies live on faces kstart+1 .. kend-1
    {
        a.wt[ijk] += - (interp2(u[ijk + 1 - kk], u[ijk + 1]) * i2x_face<TF, H>(w, ijk + 1, 1)
                      - interp2(u[ijk - kk], u[ijk]) * i2x_face<TF, H>(w, ijk, 1)) * dxi
                     - (interp2(v[ijk + jj - kk], v[ijk + jj]) * i2x_face<TF, H>(w, ijk + jj, jj)
                      - interp2(v[ijk - kk], v[ijk]) * i2x_face<TF, H>(w, ijk, jj)) * dyi
                     - (g.rhoref[k] * interp2(w[ijk], w[ijk + kk]) * i2x_wcentre<TF, H>(w, ijk, kk, k, ks, ke)
                      - g.rhoref[k - 1] * interp2(w[ijk - kk], w[ijk]) * i2x_wcentre<TF, H>(w, ijk - kk, kk, k - 1, ks, ke)) / rhoh0 * g.dzhi[k];
    }
}

template <typename TF, int H>
__global__ void __launch_bounds__(256) adv2i_s_kernel(TF* __restrict__ st, const TF* __restrict__ s, const TF* __restrict__ u,
        const TF* __restrict__ v, const TF* __restrict__ w, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ijk = i + j * jj + k * kk;
    bool okt, okb;
    const TF top = i2x_vface<TF, H>(s, ijk + kk, kk, k + 1, g.kstart, g.kend, okt), bot = i2x_vface<TF, H>(s, ijk, kk, k, g.kstart, g.kend, okb);
    const TF ft = okt ? g.rhorefh[k + 1] * w[ijk + kk] * top : TF(0);
    const TF fb = okb ? g.rhorefh[k] * w[ijk] * bot : TF(0);
    st[ijk] += - (u[ijk + 1] * i2x_face<TF, H>(s, ijk + 1, 1) - u[ijk] * i2x_face<TF, H>(s, ijk, 1)) * g.dxi
               - (v[ijk + jj] * i2x_face<TF, H>(s, ijk + jj, jj) - v[ijk] * i2x_face<TF, H>(s, ijk, jj)) * g.dyi
               - (ft - fb) / g.rhoref[k] * g.dzi[k];
}

// calc_cfl (src/advec_2i4.cxx:53-107, src/advec_2i62.cxx:59-102): the velocities interpolated to the cell centre with the scheme's
// own interpolation (2i4: w with interp2 in the lowest and the highest cell)
template <typename TF, int H>
__global__ void __launch_bounds__(256) adv2i_cfl_kernel(const TF* __restrict__ u, const TF* __restrict__ v, const TF* __restrict__ w,
        const GridDev<TF> g, double* __restrict__ out)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    TF val = TF(0);
    if (i < g.iend && j < g.jend)
    {
        const long long jj = g.icells, kk = g.ijcells;
        const long long ijk = i + j * jj + k * kk;
        val = absf(i2x_face<TF, H>(u, ijk + 1, 1)) * g.dxi + absf(i2x_face<TF, H>(v, ijk + jj, jj)) * g.dyi
            + absf(i2x_wcentre<TF, H>(w, ijk, kk, k, g.kstart, g.kend)) * g.dzi[k];
    }
    block_max_to_global<TF>(val, out);
}

} // namespace mhh
