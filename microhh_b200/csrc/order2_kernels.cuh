// mhhb200 -- plain 2nd-order schemes: Advec_2 (flux form, +-1 stencil) and Diff_2 (constant viscosity).
//
// One thread per grid point, u, v, w tendencies fused in one kernel (each of u, v, w is read once per
// launch from HBM; the +-1 neighbours are L1/L2 hits), one kernel per scalar.  These schemes are 3-5x
// cheaper per point than 2i5 + smag2, so the point-wise form already runs at the HBM roofline.
//
// Reference behaviour restated (never copied):
//   Advec_2 advec_u/v/w/s, calc_cfl   src/advec_2.cxx:48-202
//   Diff_2  diff_c / diff_w           src/diff_2.cxx:38-86  (dxidxi, dyidyi are double even in the SP build)
//   thermo_dry buoyancy               src/thermo_dry.cxx:165-179
#pragma once
#include "common.cuh"
#include "stencil_kernels.cuh"

namespace mhh {

// nu * laplacian with the reference's grouping; horizontal factors in double (src/diff_2.cxx:44-45)
template <typename TF>
__device__ __forceinline__ double diff2_term(const TF* __restrict__ a, const long long ijk, const long long jj, const long long kk,
        const TF visc, const double dxidxi, const double dyidyi, const TF dz_up, const TF dz_dn, const TF dz_c)
{
    const TF c = a[ijk];
    const double lap = (double)((a[ijk + 1] - c) - (c - a[ijk - 1])) * dxidxi
                     + (double)((a[ijk + jj] - c) - (c - a[ijk - jj])) * dyidyi
                     + (double)(((a[ijk + kk] - c) * dz_up - (c - a[ijk - kk]) * dz_dn) * dz_c);
    return (double)visc * lap;          // the caller's `+=` narrows (double)at + this to TF, as the reference's does
}

template <typename TF>
struct O2Args
{
    TF* ut; TF* vt; TF* wt;
    const TF* u; const TF* v; const TF* w;
    const TF* th;
    TF visc;
    double dxidxi, dyidyi;
};

template <typename TF, bool ADV, bool DIFF, bool BUOY>
__global__ void __launch_bounds__(256) o2_uvw_kernel(const O2Args<TF> a, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ijk = i + j * jj + k * kk;
    const TF* __restrict__ u = a.u; const TF* __restrict__ v = a.v; const TF* __restrict__ w = a.w;
    const TF dxi = g.dxi, dyi = g.dyi;
    TF tu = 0, tv = 0, tw = 0;
    const bool wrow = k > g.kstart;          // w tendencies live on faces kstart+1 .. kend-1
    if (ADV)
    {
        const TF rhoh1 = g.rhorefh[k + 1], rhoh0 = g.rhorefh[k], rho = g.rhoref[k], dzi = g.dzi[k];
        tu = - (interp2(u[ijk], u[ijk + 1]) * interp2(u[ijk], u[ijk + 1])
              - interp2(u[ijk - 1], u[ijk]) * interp2(u[ijk - 1], u[ijk])) * dxi
             - (interp2(v[ijk - 1 + jj], v[ijk + jj]) * interp2(u[ijk], u[ijk + jj])
              - interp2(v[ijk - 1], v[ijk]) * interp2(u[ijk - jj], u[ijk])) * dyi
             - (rhoh1 * interp2(w[ijk - 1 + kk], w[ijk + kk]) * interp2(u[ijk], u[ijk + kk])
              - rhoh0 * interp2(w[ijk - 1], w[ijk]) * interp2(u[ijk - kk], u[ijk])) / rho * dzi;
        tv = - (interp2(u[ijk + 1 - jj], u[ijk + 1]) * interp2(v[ijk], v[ijk + 1])
              - interp2(u[ijk - jj], u[ijk]) * interp2(v[ijk - 1], v[ijk])) * dxi
             - (interp2(v[ijk], v[ijk + jj]) * interp2(v[ijk], v[ijk + jj])
              - interp2(v[ijk - jj], v[ijk]) * interp2(v[ijk - jj], v[ijk])) * dyi
             - (rhoh1 * interp2(w[ijk - jj + kk], w[ijk + kk]) * interp2(v[ijk], v[ijk + kk])
              - rhoh0 * interp2(w[ijk - jj], w[ijk]) * interp2(v[ijk - kk], v[ijk])) / rho * dzi;
        if (wrow)
            tw = - (interp2(u[ijk + 1 - kk], u[ijk + 1]) * interp2(w[ijk], w[ijk + 1])
                  - interp2(u[ijk - kk], u[ijk]) * interp2(w[ijk - 1], w[ijk])) * dxi
                 - (interp2(v[ijk + jj - kk], v[ijk + jj]) * interp2(w[ijk], w[ijk + jj])
                  - interp2(v[ijk - kk], v[ijk]) * interp2(w[ijk - jj], w[ijk])) * dyi
                 - (g.rhoref[k] * interp2(w[ijk], w[ijk + kk]) * interp2(w[ijk], w[ijk + kk])
                  - g.rhoref[k - 1] * interp2(w[ijk - kk], w[ijk]) * interp2(w[ijk - kk], w[ijk])) / rhoh0 * g.dzhi[k];
    }
    // the reference applies thermo (buoyancy), advection and diffusion as separate "+=" in that order; every
    // partial sum below is rounded to TF exactly where the reference's stores round it
    TF ut = a.ut[ijk], vt = a.vt[ijk], wt = a.wt[ijk];
    if (BUOY && wrow) wt += TF(GRAV) / g.threfh[k] * (interp2(a.th[ijk - kk], a.th[ijk]) - g.threfh[k]);
    if (ADV) { ut += tu; vt += tv; if (wrow) wt += tw; }
    if (DIFF)
    {
        ut = (TF)((double)ut + diff2_term<TF>(u, ijk, jj, kk, a.visc, a.dxidxi, a.dyidyi, g.dzhi[k + 1], g.dzhi[k], g.dzi[k]));
        vt = (TF)((double)vt + diff2_term<TF>(v, ijk, jj, kk, a.visc, a.dxidxi, a.dyidyi, g.dzhi[k + 1], g.dzhi[k], g.dzi[k]));
        if (wrow) wt = (TF)((double)wt + diff2_term<TF>(w, ijk, jj, kk, a.visc, a.dxidxi, a.dyidyi, g.dzi[k], g.dzi[k - 1], g.dzhi[k]));
    }
    a.ut[ijk] = ut; a.vt[ijk] = vt;
    if (wrow && (ADV || DIFF || BUOY)) a.wt[ijk] = wt;
}

template <typename TF>
struct O2ScalArgs
{
    TF* st;
    const TF* s; const TF* u; const TF* v; const TF* w;
    TF visc;
    double dxidxi, dyidyi;
};

template <typename TF, bool ADV, bool DIFF>
__global__ void __launch_bounds__(256) o2_s_kernel(const O2ScalArgs<TF> a, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ijk = i + j * jj + k * kk;
    const TF* __restrict__ s = a.s;
    TF st = a.st[ijk];
    if (ADV)
    {
        st += - (a.u[ijk + 1] * interp2(s[ijk], s[ijk + 1]) - a.u[ijk] * interp2(s[ijk - 1], s[ijk])) * g.dxi
              - (a.v[ijk + jj] * interp2(s[ijk], s[ijk + jj]) - a.v[ijk] * interp2(s[ijk - jj], s[ijk])) * g.dyi
              - (g.rhorefh[k + 1] * a.w[ijk + kk] * interp2(s[ijk], s[ijk + kk])
               - g.rhorefh[k] * a.w[ijk] * interp2(s[ijk - kk], s[ijk])) / g.rhoref[k] * g.dzi[k];
    }
    if (DIFF) st = (TF)((double)st + diff2_term<TF>(s, ijk, jj, kk, a.visc, a.dxidxi, a.dyidyi, g.dzhi[k + 1], g.dzhi[k], g.dzi[k]));
    a.st[ijk] = st;
}

// Advec_2 calc_cfl (src/advec_2.cxx:50-76): max of |u_c| dxi + |v_c| dyi + |w_c| dzi
template <typename TF>
__global__ void __launch_bounds__(256) o2_cfl_kernel(const TF* __restrict__ u, const TF* __restrict__ v, const TF* __restrict__ w,
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
        val = absf(interp2(u[ijk], u[ijk + 1])) * g.dxi + absf(interp2(v[ijk], v[ijk + jj])) * g.dyi
            + absf(interp2(w[ijk], w[ijk + kk])) * g.dzi[k];
    }
    block_max_to_global<TF>(val, out);
}

} // namespace mhh
