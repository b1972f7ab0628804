// mhhb200 -- Monin-Obukhov surface model (Boundary_surface<TF>::exec with constant z0 and the lookup solver).
//
// Reference behaviour restated (never copied): src/boundary_surface.cxx:55-340 (stability, stability_neutral, surfm, surfs),
// :836-990 (call order); include/boundary_surface_kernels.h:78-330 (prepare_lut, calc_dutot, find_zL, calc_duvdz_mo,
// calc_dbdz_mo); include/monin_obukhov.h (Wilson 2001 / IFS stability functions); Thermo_dry's surface buoyancy
// (src/thermo_dry.cxx:133-162, 700-784).
//
// Everything here is 2-D (one horizontal plane): a sub-step spends ~10 us in these kernels.  They exist so that a
// self-driven multi-step LES stays resident on the device: exec_viscosity -> [this] -> set_ghost_cells ->
// tendencies is the reference's order (src/model.cxx:375-401).  Four launches:
//   surface_dutot_kernel      filtered wind speed difference at the first level (interior), then the cyclic fill
//   surface_stability_kernel  Obukhov length from the z/L lookup table, friction velocity (all cells incl. ghosts)
//   surface_flux_kernel       momentum fluxes (interior; cyclic fill afterwards), MO gradients dudz/dvdz/dbdz
//   surface_values_kernel     u/v gradients and the scalars' surface value or flux + gradient (all cells)
#pragma once
#include "common.cuh"

namespace mhh {

constexpr int SURF_NLUT = 10000;      // nzL_lut (include/boundary.h:56)

template <typename TF> __device__ __forceinline__ TF powf_(TF a, TF b);
template <> __device__ __forceinline__ double powf_<double>(double a, double b) { return pow(a, b); }
template <> __device__ __forceinline__ float powf_<float>(float a, float b) { return powf(a, b); }
template <typename TF> __device__ __forceinline__ TF logf_(TF a);
template <> __device__ __forceinline__ double logf_<double>(double a) { return log(a); }
template <> __device__ __forceinline__ float logf_<float>(float a) { return logf(a); }
template <typename TF> __device__ __forceinline__ TF expm_(TF a);
template <> __device__ __forceinline__ double expm_<double>(double a) { return exp(a); }
template <> __device__ __forceinline__ float expm_<float>(float a) { return expf(a); }

namespace most {
template <typename TF> __host__ __device__ inline TF kappa() { return TF(0.4); }
template <typename TF> __device__ __forceinline__ TF phim_unstable(TF zeta)
{ return powf_<TF>(TF(1.) + TF(3.6) * powf_<TF>(absf(zeta), TF(2. / 3.)), TF(-1. / 2.)); }
template <typename TF> __device__ __forceinline__ TF phih_unstable(TF zeta)
{ return powf_<TF>(TF(1.) + TF(7.9) * powf_<TF>(absf(zeta), TF(2. / 3.)), TF(-1. / 2.)); }
template <typename TF> __device__ __forceinline__ TF phim(TF zeta)
{ return zeta <= TF(0.) ? phim_unstable(zeta) : TF(1) + TF(5) * zeta; }
template <typename TF> __device__ __forceinline__ TF phih(TF zeta)
{ return zeta <= TF(0.) ? phih_unstable(zeta) : pow2(TF(1) + TF(4) * zeta); }
template <typename TF> __device__ __forceinline__ TF psim_unstable(TF zeta)
{ return TF(3.) * logf_<TF>((TF(1.) + TF(1.) / phim_unstable(zeta)) / TF(2.)); }
template <typename TF> __device__ __forceinline__ TF psih_unstable(TF zeta)
{ return TF(3.) * logf_<TF>((TF(1.) + TF(1.) / phih_unstable(zeta)) / TF(2.)); }
template <typename TF> __device__ __forceinline__ TF psim_stable(TF zeta)
{
    const TF a = TF(1), b = TF(2) / TF(3), c = TF(5), d = TF(0.35);
    return -b * (zeta - (c / d)) * expm_<TF>(-d * zeta) - a * zeta - (b * c) / d;
}
template <typename TF> __device__ __forceinline__ TF psih_stable(TF zeta)
{
    const TF a = TF(1), b = TF(2) / TF(3), c = TF(5), d = TF(0.35);
    return -b * (zeta - (c / d)) * expm_<TF>(-d * zeta) - powf_<TF>(TF(1) + b * a * zeta, TF(1.5)) - (b * c) / d + TF(1);
}
template <typename TF> __device__ __forceinline__ TF fm(TF zsl, TF z0m, TF L)
{
    return (L <= TF(0.)) ? kappa<TF>() / (logf_<TF>(zsl / z0m) - psim_unstable(zsl / L) + psim_unstable(z0m / L))
                         : kappa<TF>() / (logf_<TF>(zsl / z0m) - psim_stable(zsl / L) + psim_stable(z0m / L));
}
template <typename TF> __device__ __forceinline__ TF fh(TF zsl, TF z0h, TF L)
{
    return (L <= TF(0.)) ? kappa<TF>() / (logf_<TF>(zsl / z0h) - psih_unstable(zsl / L) + psih_unstable(z0h / L))
                         : kappa<TF>() / (logf_<TF>(zsl / z0h) - psih_stable(zsl / L) + psih_stable(z0h / L));
}
} // namespace most

template <typename TF>
struct SurfArgs
{
    // state of the surface model (2-D, ijcells)
    TF* ustar; TF* obuk; int* nobuk; TF* dutot;
    const TF* z0m; const TF* z0h;
    const float* zL_sl; const float* f_sl;       // lookup table (SURF_NLUT)
    // first-level fields and their 2-D companions
    const TF* u; const TF* v;
    const TF* ubot; const TF* vbot;
    TF* ufluxbot; TF* vfluxbot; TF* ugradbot; TF* vgradbot;
    TF* dudz; TF* dvdz; TF* dbdz;
    int ns;
    const TF* s[MAX_SCALARS]; TF* sbot[MAX_SCALARS]; TF* sgradbot[MAX_SCALARS]; TF* sfluxbot[MAX_SCALARS];
    int sbc[MAX_SCALARS];                         // 0 Dirichlet, 2 Flux, anything else: untouched
    int mbcbot, thermobc, neutral;                // 0 Dirichlet (no-slip), 2 Flux, 3 Ustar
    TF zsl, gthref, gthrefh, thref, threfh;       // z[kstart]; g/thref[kstart], g/threfh[kstart], thref[kstart], threfh[kstart]
};

// bsk::calc_dutot: 3 x 4 filtered wind at the scalar location minus the surface value, floored at 0.1 m/s (interior only)
template <typename TF>
__global__ void surface_dutot_kernel(const SurfArgs<TF> a, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells;
    const long long ij = i + j * jj;
    const TF* __restrict__ u = a.u + g.kstart * g.ijcells + ij;
    const TF* __restrict__ v = a.v + g.kstart * g.ijcells + ij;
    const TF h = TF(0.5);
    const TF uf = TF(1. / 9) *
        ( h * u[-1 - jj] + u[-jj] + u[1 - jj] + h * u[2 - jj]
        + h * u[-1     ] + u[0  ] + u[1     ] + h * u[2     ]
        + h * u[-1 + jj] + u[jj ] + u[1 + jj] + h * u[2 + jj] );
    const TF vf = TF(1. / 9) *
        ( h * v[-1 - jj] + v[-1] + v[-1 + jj] + h * v[-1 + 2 * jj]
        + h * v[   - jj] + v[0 ] + v[     jj] + h * v[     2 * jj]
        + h * v[ 1 - jj] + v[1 ] + v[ 1 + jj] + h * v[ 1 + 2 * jj] );
    const TF du2 = pow2(uf - h * (a.ubot[ij] + a.ubot[ij + 1])) + pow2(vf - h * (a.vbot[ij] + a.vbot[ij + jj]));
    const TF d = sqrtf_(du2);
    a.dutot[ij] = d > TF(1.e-1) ? d : TF(1.e-1);
}

// bsk::find_zL: bracket search from the previous index at float accuracy, then linear interpolation
template <typename TF>
__device__ __forceinline__ TF surf_find_zL(const float* __restrict__ zL, const float* __restrict__ f, int& n, const float Ri)
{
    if ((f[n] - Ri) > 0.f) { while (n > 0 && (f[n - 1] - Ri) > 0.f) --n; }
    else { while ((f[n] - Ri) < 0.f && n < SURF_NLUT - 1) ++n; }
    return (n == 0 || n == SURF_NLUT - 1) ? (TF)zL[n] : (TF)(zL[n - 1] + (Ri - f[n - 1]) / (f[n] - f[n - 1]) * (zL[n] - zL[n - 1]));
}

// stability / stability_neutral over ALL cells (the reference loops 0..icells, 0..jcells: dutot has been cyclic-filled)
template <typename TF>
__global__ void surface_stability_kernel(const SurfArgs<TF> a, const GridDev<TF> g)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.icells || j >= g.jcells) return;
    const long long ij = i + (long long)j * g.icells;
    const TF zsl = a.zsl;
    if (a.neutral)
    {
        if (a.mbcbot == 3) { if (i >= g.istart && i < g.iend && j >= g.jstart && j < g.jend) a.obuk[ij] = -TF(1.e9); return; }
        a.obuk[ij] = -TF(1.e9);
        a.ustar[ij] = a.dutot[ij] * most::fm<TF>(zsl, a.z0m[ij], -TF(1.e9));
        return;
    }
    const TF* th = a.s[0];
    const TF bfluxbot = a.gthrefh * a.sfluxbot[0][ij];
    if (a.mbcbot == 3 && a.thermobc == 2)
    {
        const TF us = a.ustar[ij];
        a.obuk[ij] = -(us * us * us) / (most::kappa<TF>() * bfluxbot);
        return;
    }
    const TF du = a.dutot[ij];
    int n = a.nobuk[ij];
    n = n < 0 ? 0 : (n > SURF_NLUT - 1 ? SURF_NLUT - 1 : n);
    float Ri;
    if (a.thermobc == 2) Ri = (float)(-most::kappa<TF>() * bfluxbot * zsl / (du * du * du));
    else
    {
        const TF bbot = a.gthrefh * (a.sbot[0][ij] - a.threfh);
        const TF b = a.gthref * (th[ij + g.kstart * g.ijcells] - a.thref);
        const TF db_ref = a.gthref * (a.thref - a.threfh);
        const TF db = b - bbot + db_ref;
        Ri = (float)(most::kappa<TF>() * db * zsl / (du * du));
    }
    const TF zL = surf_find_zL<TF>(a.zL_sl, a.f_sl, n, Ri);
    const TF L = zsl / zL;
    a.nobuk[ij] = n;
    a.obuk[ij] = L;
    a.ustar[ij] = du * most::fm<TF>(zsl, a.z0m[ij], L);
}

// surfm (no-slip: fluxes from the interpolated stability function), calc_duvdz_mo, calc_dbdz_mo: interior
template <typename TF>
__global__ void surface_flux_kernel(const SurfArgs<TF> a, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells;
    const long long ij = i + j * jj;
    const long long ijk = ij + g.kstart * g.ijcells;
    const TF zsl = a.zsl;
    const TF fm_c = most::fm<TF>(zsl, a.z0m[ij], a.obuk[ij]);
    const TF sf_c = a.ustar[ij] * fm_c;
    const TF sf_w = a.ustar[ij - 1] * most::fm<TF>(zsl, a.z0m[ij - 1], a.obuk[ij - 1]);
    const TF sf_s = a.ustar[ij - jj] * most::fm<TF>(zsl, a.z0m[ij - jj], a.obuk[ij - jj]);
    a.ufluxbot[ij] = -(a.u[ijk] - a.ubot[ij]) * TF(0.5) * (sf_w + sf_c);
    a.vfluxbot[ij] = -(a.v[ijk] - a.vbot[ij]) * TF(0.5) * (sf_s + sf_c);
    // MO gradients at the scalar location
    const TF du_c = TF(0.5) * ((a.u[ijk] - a.ubot[ij]) + (a.u[ijk + 1] - a.ubot[ij + 1]));
    const TF dv_c = TF(0.5) * ((a.v[ijk] - a.vbot[ij]) + (a.v[ijk + jj] - a.vbot[ij + jj]));
    const TF uflux = -du_c * a.ustar[ij] * fm_c;
    const TF vflux = -dv_c * a.ustar[ij] * fm_c;
    const TF phim = most::phim<TF>(zsl / a.obuk[ij]);
    a.dudz[ij] = -uflux / (most::kappa<TF>() * zsl * a.ustar[ij]) * phim;
    a.dvdz[ij] = -vflux / (most::kappa<TF>() * zsl * a.ustar[ij]) * phim;
    if (!a.neutral)
    {
        const TF bfluxbot = a.gthrefh * a.sfluxbot[0][ij];
        a.dbdz[ij] = -bfluxbot / (most::kappa<TF>() * zsl * a.ustar[ij]) * most::phih<TF>(zsl / a.obuk[ij]);
    }
}

// surfm's linearly interpolated gradients and surfs for every scalar: all cells
template <typename TF>
__global__ void surface_values_kernel(const SurfArgs<TF> a, const GridDev<TF> g)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.icells || j >= g.jcells) return;
    const long long ij = i + (long long)j * g.icells;
    const long long ijk = ij + g.kstart * g.ijcells;
    const TF zsl = a.zsl;
    a.ugradbot[ij] = (a.u[ijk] - a.ubot[ij]) / zsl;
    a.vgradbot[ij] = (a.v[ijk] - a.vbot[ij]) / zsl;
    if (a.ns == 0) return;
    const TF fh = most::fh<TF>(zsl, a.z0h[ij], a.obuk[ij]);
    const TF us = a.ustar[ij];
    for (int n = 0; n < a.ns; ++n)
    {
        const TF var = a.s[n][ijk];
        if (a.sbc[n] == 0)
        {
            const TF vb = a.sbot[n][ij];
            a.sfluxbot[n][ij] = -(var - vb) * us * fh;
            a.sgradbot[n][ij] = (var - vb) / zsl;
        }
        else if (a.sbc[n] == 2)
        {
            const TF vb = a.sfluxbot[n][ij] / (us * fh) + var;
            a.sbot[n][ij] = vb;
            a.sgradbot[n][ij] = (var - vb) / zsl;
        }
    }
}

} // namespace mhh
