// mhhb200 -- Deardorff (1980) SGS-TKE closure, Diff_tke2<TF> of the reference (src/diff_tke2.cxx), and the Limiter's
// tendency limiter (src/limiter.cxx:35-59).
//
// Diff_tke2::exec_viscosity (src/diff_tke2.cxx:799-983) is seven passes over the grid in the reference (strain^2, N2, evisc,
// eviscs, buoyancy / dissipation / shear tendencies of sgstke, through two scratch fields); here it is ONE kernel: every point
// forms its strain^2 (Diff_kernels::calc_strain2 with the surface model's gradients at the lowest level), its N2 (array, or
// from th as Thermo_dry does), the two eddy viscosities and the three sources of sgstke in registers.  Algorithmic traffic:
// read u, v, w, sgstke, th; write evisc, eviscs; read-modify-write the sgstke tendency = 9 array passes (reference: 27).
// Diff_tke2::exec itself is Diff_kernels::diff_u / diff_v / diff_w / diff_c with tPr = 1 -- the same fused tendency
// kernels as Diff_smag2, with the eddy viscosity chosen per scalar (host_tend.cu).
#pragma once
#include "common.cuh"
#include "stencil_kernels.cuh"

namespace mhh {

#define MHH_SGSTKE_MIN 1.e-7     // Constants::sgstke_min (include/constants.h:59)

template <typename TF>
struct Tke2Args
{
    TF* evisc; TF* eviscs;        // eviscs: only with buoyancy
    TF* st;                       // tendency of sgstke
    const TF* e;                  // sgstke
    const TF* u; const TF* v; const TF* w;
    const TF* n2;                 // n2mode 0
    const TF* th;                 // n2mode 1
    const TF* dudz; const TF* dvdz; const TF* dbdz; const TF* z0m;
    TF cn, cm, ch1, ch2, ce1, ce2;
    int mason, buoy, n2mode;
};

// Mason's wall correction with n = 2: 1 / fac^2 = 1 / l^2 + 1 / (kappa (z + z0m))^2
template <typename TF>
__device__ __forceinline__ TF tke2_mason(const TF l, const TF zz)
{
    const TF t = TF(KAPPA) * zz;
    return sqrtf_(TF(1.) / (TF(1.) / (l * l) + TF(1.) / (t * t)));
}

template <typename TF>
__global__ void __launch_bounds__(256) tke2_visc_kernel(const Tke2Args<TF> a, const GridDev<TF> g, const TF* __restrict__ mlen0)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ij = i + j * jj;
    const long long ijk = ij + k * kk;
    const bool bottom = (k == g.kstart);
    // Surface_model::Enabled throughout ("Resolved wall not supported in Deardorff SGSm.", src/diff_tke2.cxx:97)
    const TF s2 = strain2_point<TF>(a.u, a.v, a.w, ijk, 1, jj, kk, g.dxi, g.dyi, g.dzi[k], g.dzhi[k], g.dzhi[k + 1],
                                    bottom, bottom ? a.dudz[ij] : TF(0), bottom ? a.dvdz[ij] : TF(0));
    const TF e = a.e[ijk];
    const TF se = sqrtf_(e);
    const TF m0 = mlen0[k];
    const TF z0 = a.z0m[ij];
    TF st = a.st[ijk];
    if (!a.buoy)
    {
        // calc_evisc_neutral (:73-134; the n = 2 Mason branch uses z[kstart] at every level, :118) and
        // sgstke_diss_tend_neutral (:471-511, z[k])
        const TF fac_e = a.mason ? tke2_mason<TF>(m0, g.z[g.kstart] + z0) : m0;
        const TF fac_d = a.mason ? tke2_mason<TF>(m0, g.z[k] + z0) : m0;
        const TF ev = a.cm * fac_e * se;
        a.evisc[ijk] = ev;
        st -= (a.ce1 + a.ce2 * fac_d / m0) * (e * se) / fac_d;
        st += ev * s2;
    }
    else
    {
        TF n2;
        if (bottom) n2 = a.dbdz[ij];
        else if (a.n2mode == 0) n2 = a.n2[ijk];
        else n2 = TF(GRAV) / g.thref[k] * TF(0.5) * (a.th[ijk + kk] - a.th[ijk - kk]) * g.dzi[k];
        // calc_evisc / calc_evisc_heat (:136-332): only if stably stratified, adapt the length scale
        const TF mlen = n2 > TF(0) ? a.cn * sqrtf_(e / n2) : m0;
        TF fac = mlen < m0 ? mlen : m0;
        if (a.mason) fac = tke2_mason<TF>(fac, g.z[k] + z0);
        const TF ev = a.cm * fac * se;
        const TF evh = (a.ch1 + a.ch2 * fac / m0) * ev;
        a.evisc[ijk] = ev;
        a.eviscs[ijk] = evh;
        // sgstke_buoy_tend (:359-393), sgstke_diss_tend (:395-469; same length scale), in the reference's order
        st -= evh * n2;
        st -= (a.ce1 + a.ce2 * fac / m0) * (e * se) / fac;
        st += ev * s2;                                     // sgstke_shear_tend (:334-357)
    }
    a.st[ijk] = st;
}

// enforce_min_sgstke (src/diff_tke2.cxx:48-71), interior only; the caller follows with the cyclic fill
template <typename TF>
__global__ void __launch_bounds__(256) tke2_min_kernel(TF* __restrict__ e, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long ijk = i + (long long)j * g.icells + (long long)k * g.ijcells;
    const TF v = e[ijk];
    e[ijk] = v > TF(MHH_SGSTKE_MIN) ? v : TF(MHH_SGSTKE_MIN);
}

// tendency_limiter (src/limiter.cxx:35-59): a source that keeps a + dt * at at or above min_value
template <typename TF>
__global__ void __launch_bounds__(256) limiter_kernel(TF* __restrict__ at, const TF* __restrict__ a, const TF min_value, const TF dt,
        const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long ijk = i + (long long)j * g.icells + (long long)k * g.ijcells;
    const TF dti = TF(1.) / dt;
    const TF t = at[ijk];
    const TF a_new = a[ijk] + dt * t;
    if (a_new < min_value) at[ijk] = t + (-a_new + min_value) * dti;
}

} // namespace mhh
