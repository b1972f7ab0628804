// mhhb200 -- point-wise (one thread per grid point) stencil kernels.
//
// These are the "plain" versions of every stencil stage of the hot path: simple, coalesced
// along x, relying on L1/L2 for neighbour reuse.  They define the arithmetic (flux form of the
// reference's expressions) and serve the individual Advec/Diff/Pres/Boundary entry points.
// The fused z-marching tile kernels in tile_kernels.cuh are the fast path of mhh_dycore_substep.
//
// Reference behaviour restated (never copied):
//   Boundary_cyclic::exec           src/boundary_cyclic.cxx:369-443
//   calc_ghost_cells_{bot,top}_2nd  src/boundary.cxx:700-772
//   Advec_2i5 advec_u/v/w/s         src/advec_2i5.cxx:151-728, calc_cfl :60-148
//   Diff_kernels::calc_strain2 etc. include/diff_kernels.h:34-511
//   calc_evisc                      src/diff_smag2.cxx:148-269
//   calc_N2 / buoyancy_tend_2nd     src/thermo_dry.cxx:66-78, 165-179
//   Pres_2::input/output/divergence src/pres_2.cxx:155-196, 364-422
//   rk3                             src/timeloop.cxx:250-286
#pragma once
#include "common.cuh"

namespace mhh {

constexpr double DSMALL = 1.e-9;   // Constants::dsmall
constexpr double KAPPA  = 0.4;     // Constants::kappa
constexpr double GRAV   = 9.81;    // Constants::grav

// ------------------------------------------------------------------------------------------
// Periodic ghost-cell fill.  One thread per ghost cell; the source is the wrapped interior
// cell, which reproduces the reference's "x first, then y" ordering (corners included)
// without any ordering dependency, so both directions go in ONE launch.
// edges: 0 = east-west only, 1 = north-south only, 2 = both.
// ------------------------------------------------------------------------------------------
template <typename TF>
__global__ void cyclic_kernel(TF* __restrict__ a, const GridDev<TF> g, const int edges, const int nk)
{
    // region 0: x strips (2*igc wide); region 1: y strips (2*jgc tall)
    const int region = blockIdx.z;
    const int k = blockIdx.y;   // 0..nk-1 (nk = kcells for 3-D fields, 1 for 2-D slices)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const long long kk = g.ijcells;
    const bool two_d_run = (g.jtot == 1);

    if (region == 0)
    {
        if (edges == 1) return;
        // x strips over all rows j (ghost rows too, as the reference does)
        const int w = 2 * g.igc;
        if (t >= w * g.jcells) return;
        const int j = t / w;
        const int s = t - j * w;
        const int i = s < g.igc ? s : g.iend + (s - g.igc);
        const int isrc = s < g.igc ? i + g.imax : i - g.imax;
        int jsrc = j;
        if (edges == 2 && (j < g.jstart || j >= g.jend))
            return;   // corner cells belong to region 1 when both edges are filled
        a[i + (long long)j * g.icells + k * kk] = a[isrc + (long long)jsrc * g.icells + k * kk];
    }
    else
    {
        if (edges == 0) return;
        const int h = 2 * g.jgc;
        if (t >= h * g.icells) return;
        const int s = t / g.icells;
        const int i = t - s * g.icells;
        const int j = s < g.jgc ? s : g.jend + (s - g.jgc);
        int jsrc = s < g.jgc ? j + g.jmax : j - g.jmax;
        int isrc = i;
        if (edges == 2)
        {
            if (i < g.istart) isrc = i + g.imax;
            else if (i >= g.iend) isrc = i - g.imax;
        }
        if (two_d_run)
        {
            // jtot == 1: replicate the single row, interior levels only (3-D), all for 2-D slices
            const bool interior_k = (nk == 1) || (k >= g.kstart && k < g.kend);
            if (interior_k) jsrc = g.jstart;
            else
            {
                // ghost levels: only the x fill applies (to every row, ghost rows included)
                if (edges != 2 || isrc == i) return;
                jsrc = j;
            }
        }
        a[i + (long long)j * g.icells + k * kk] = a[isrc + (long long)jsrc * g.icells + k * kk];
    }
}

// ------------------------------------------------------------------------------------------
// Vertical ghost cells, 2nd order.  bc: 0 Dirichlet (value given), 1 Neumann/flux (gradient given)
// ------------------------------------------------------------------------------------------
template <typename TF>
__global__ void ghost_cells_2nd_kernel(TF* __restrict__ a, const GridDev<TF> g,
        const int bcbot, const TF* __restrict__ abot, const TF* __restrict__ agradbot,
        const int bctop, const TF* __restrict__ atop, const TF* __restrict__ agradtop)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.icells || j >= g.jcells) return;
    const long long ij = i + (long long)j * g.icells;
    const long long kk = g.ijcells;
    if (bcbot >= 0)
    {
        const long long ijk = ij + g.kstart * kk;
        if (bcbot == 0) a[ijk - kk] = TF(2.) * abot[ij] - a[ijk];
        else            a[ijk - kk] = -agradbot[ij] * g.dzh[g.kstart] + a[ijk];
    }
    if (bctop >= 0)
    {
        const long long ijk = ij + (g.kend - 1) * kk;
        if (bctop == 0) a[ijk + kk] = TF(2.) * atop[ij] - a[ijk];
        else            a[ijk + kk] = agradtop[ij] * g.dzh[g.kend] + a[ijk];
    }
}

// ------------------------------------------------------------------------------------------
// Eddy viscosity: strain^2 (Diff_kernels::calc_strain2) + N2 (Thermo_dry calc_N2) + Smagorinsky-
// Lilly with Mason wall damping and stability correction (calc_evisc), fused: u,v,w,th -> evisc.
// n2mode: 0 = N2 array supplied, 1 = N2 from th (thermo_dry).
// ------------------------------------------------------------------------------------------
template <typename TF>
struct EviscArgs
{
    TF* evisc;
    const TF* u; const TF* v; const TF* w;
    const TF* n2;         // n2mode 0
    const TF* th;         // n2mode 1
    const TF* dudz; const TF* dvdz; const TF* dbdz; const TF* z0m;   // 2-D, surface model only
    TF cs, tPr;
    int surface, mason, n2mode;
};

template <typename TF>
__device__ __forceinline__ TF strain2_point(const TF* __restrict__ u, const TF* __restrict__ v, const TF* __restrict__ w,
        const long long ijk, const int ii, const long long jj, const long long kk,
        const TF dxi, const TF dyi, const TF dzi_k, const TF dzhi_k, const TF dzhi_kp,
        const bool bottom_mo, const TF dudz_b, const TF dvdz_b)
{
    const TF e = TF(0.125);
    TF s = pow2((u[ijk + ii] - u[ijk]) * dxi)
         + pow2((v[ijk + jj] - v[ijk]) * dyi)
         + pow2((w[ijk + kk] - w[ijk]) * dzi_k);
    s += e * pow2((u[ijk          ] - u[ijk      - jj]) * dyi + (v[ijk          ] - v[ijk - ii     ]) * dxi);
    s += e * pow2((u[ijk + ii     ] - u[ijk + ii - jj]) * dyi + (v[ijk + ii     ] - v[ijk          ]) * dxi);
    s += e * pow2((u[ijk      + jj] - u[ijk          ]) * dyi + (v[ijk      + jj] - v[ijk - ii + jj]) * dxi);
    s += e * pow2((u[ijk + ii + jj] - u[ijk + ii     ]) * dyi + (v[ijk + ii + jj] - v[ijk      + jj]) * dxi);
    if (bottom_mo)
    {
        s += TF(0.5) * pow2(dudz_b);
        s += e * pow2((w[ijk          ] - w[ijk - ii     ]) * dxi);
        s += e * pow2((w[ijk + ii     ] - w[ijk          ]) * dxi);
        s += e * pow2((w[ijk      + kk] - w[ijk - ii + kk]) * dxi);
        s += e * pow2((w[ijk + ii + kk] - w[ijk      + kk]) * dxi);
        s += TF(0.5) * pow2(dvdz_b);
        s += e * pow2((w[ijk          ] - w[ijk - jj     ]) * dyi);
        s += e * pow2((w[ijk + jj     ] - w[ijk          ]) * dyi);
        s += e * pow2((w[ijk      + kk] - w[ijk - jj + kk]) * dyi);
        s += e * pow2((w[ijk + jj + kk] - w[ijk      + kk]) * dyi);
    }
    else
    {
        s += e * pow2((u[ijk          ] - u[ijk      - kk]) * dzhi_k  + (w[ijk          ] - w[ijk - ii     ]) * dxi);
        s += e * pow2((u[ijk + ii     ] - u[ijk + ii - kk]) * dzhi_k  + (w[ijk + ii     ] - w[ijk          ]) * dxi);
        s += e * pow2((u[ijk      + kk] - u[ijk          ]) * dzhi_kp + (w[ijk      + kk] - w[ijk - ii + kk]) * dxi);
        s += e * pow2((u[ijk + ii + kk] - u[ijk + ii     ]) * dzhi_kp + (w[ijk + ii + kk] - w[ijk      + kk]) * dxi);
        s += e * pow2((v[ijk          ] - v[ijk      - kk]) * dzhi_k  + (w[ijk          ] - w[ijk - jj     ]) * dyi);
        s += e * pow2((v[ijk + jj     ] - v[ijk + jj - kk]) * dzhi_k  + (w[ijk + jj     ] - w[ijk          ]) * dyi);
        s += e * pow2((v[ijk      + kk] - v[ijk          ]) * dzhi_kp + (w[ijk      + kk] - w[ijk - jj + kk]) * dyi);
        s += e * pow2((v[ijk + jj + kk] - v[ijk + jj     ]) * dzhi_kp + (w[ijk + jj + kk] - w[ijk      + kk]) * dyi);
    }
    return (TF)((double)(TF(2.) * s) + DSMALL);
}

// Smagorinsky mixing length squared.  mlen0 = cs*(dx*dy*dz)^(1/3); Mason (n=2):
// mlen^2 = 1/(1/mlen0^2 + 1/(kappa*(z+z0m))^2).
template <typename TF>
__device__ __forceinline__ TF mlen2_of(const TF mlen0, const bool mason, const TF zk, const TF z0)
{
    if (!mason) return mlen0 * mlen0;
    const TF t = TF(KAPPA) * (zk + z0);
    const TF m = sqrtf_(TF(1.) / (TF(1.) / (mlen0 * mlen0) + TF(1.) / (t * t)));
    return m * m;
}

template <typename TF>
__global__ void evisc_kernel(const EviscArgs<TF> a, const GridDev<TF> g, const TF* __restrict__ mlen0)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ij = i + j * jj;
    const long long ijk = ij + k * kk;
    const bool bottom_mo = a.surface && (k == g.kstart);
    const TF s2 = strain2_point<TF>(a.u, a.v, a.w, ijk, 1, jj, kk, g.dxi, g.dyi, g.dzi[k], g.dzhi[k], g.dzhi[k + 1],
                                    bottom_mo, bottom_mo ? a.dudz[ij] : TF(0), bottom_mo ? a.dvdz[ij] : TF(0));
    TF n2;
    if (bottom_mo) n2 = a.dbdz[ij];
    else if (a.n2mode == 0) n2 = a.n2[ijk];
    else n2 = TF(GRAV) / g.thref[k] * TF(0.5) * (a.th[ijk + kk] - a.th[ijk - kk]) * g.dzi[k];
    TF rit = n2 / s2 / a.tPr;
    rit = rit < TF(1. - DSMALL) ? rit : TF(1. - DSMALL);
    const TF ml0 = a.cs * mlen0[k];
    const TF m2 = a.surface ? mlen2_of<TF>(ml0, a.mason != 0, g.z[k], a.z0m[ij]) : ml0 * ml0;
    a.evisc[ijk] = m2 * sqrtf_(s2) * sqrtf_(TF(1.) - rit);
}

// Neutral eddy viscosity (no thermo; reference calc_evisc_neutral, src/diff_smag2.cxx:47-146): K = mlen^2 sqrt(S^2).
// Surface model: Mason wall correction with n = 1, mlen = 1 / (1/mlen0 + 1/(kappa (z + z0m))).  Resolved walls: van Driest
// damping, mlen = fac * mlen0 with fac = min over both walls of 1 - exp(-(distance * u_tau) / (26 nu)),
// u_tau = |nu dU/dz at the wall|^(1/2).
template <typename TF> __device__ __forceinline__ TF expf_(TF a);
template <> __device__ __forceinline__ double expf_<double>(double a) { return exp(a); }
template <> __device__ __forceinline__ float expf_<float>(float a) { return expf(a); }

template <typename TF>
__global__ void evisc_neutral_kernel(const EviscArgs<TF> a, const GridDev<TF> g, const TF* __restrict__ mlen0, const TF visc)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ij = i + j * jj;
    const long long ijk = ij + k * kk;
    const bool bottom_mo = a.surface && (k == g.kstart);
    const TF s2 = strain2_point<TF>(a.u, a.v, a.w, ijk, 1, jj, kk, g.dxi, g.dyi, g.dzi[k], g.dzhi[k], g.dzhi[k + 1],
                                    bottom_mo, bottom_mo ? a.dudz[ij] : TF(0), bottom_mo ? a.dvdz[ij] : TF(0));
    const TF ml0 = a.cs * mlen0[k];
    TF mlen;
    if (a.surface)
        mlen = a.mason ? TF(1.) / (TF(1.) / ml0 + TF(1.) / (TF(KAPPA) * (g.z[k] + a.z0m[ij]))) : ml0;
    else
    {
        const long long ib = ij + g.kstart * kk, it = ij + g.kend * kk;
        const TF tb = pow2(visc * (a.u[ib] - a.u[ib - kk]) * g.dzhi[g.kstart]) + pow2(visc * (a.v[ib] - a.v[ib - kk]) * g.dzhi[g.kstart]);
        const TF tt = pow2(visc * (a.u[it] - a.u[it - kk]) * g.dzhi[g.kend]) + pow2(visc * (a.v[it] - a.v[it - kk]) * g.dzhi[g.kend]);
        const TF utau_b = sqrtf_(sqrtf_(tb)), utau_t = sqrtf_(sqrtf_(tt));          // pow(x, 1/4)
        const TF fac_b = TF(1.) - expf_<TF>(-(g.z[k] * utau_b) / (TF(26.) * visc));
        const TF fac_t = TF(1.) - expf_<TF>(-((g.zsize - g.z[k]) * utau_t) / (TF(26.) * visc));
        mlen = (fac_b < fac_t ? fac_b : fac_t) * ml0;
    }
    a.evisc[ijk] = mlen * mlen * sqrtf_(s2);
}

// Resolved-wall variant: mirror evisc over bottom and top walls (src/diff_smag2.cxx:195-207).
template <typename TF>
__global__ void evisc_mirror_kernel(TF* __restrict__ evisc, const GridDev<TF> g)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.icells || j >= g.jcells) return;
    const long long ij = i + (long long)j * g.icells;
    const long long kk = g.ijcells;
    evisc[ij + (g.kstart - 1) * kk] = evisc[ij + g.kstart * kk];
    evisc[ij + g.kend * kk] = evisc[ij + (g.kend - 1) * kk];
}

// ------------------------------------------------------------------------------------------
// Momentum tendencies: Advec_2i5 u/v/w + Diff_smag2 u/v/w (+ thermo_dry buoyancy), flux form.
// ------------------------------------------------------------------------------------------
template <typename TF>
struct MomArgs
{
    TF* ut; TF* vt; TF* wt;
    const TF* u; const TF* v; const TF* w;
    const TF* evisc;
    const TF* th;                       // buoyancy (thermo_dry) when BUOY
    const TF* u_fluxbot; const TF* u_fluxtop; const TF* v_fluxbot; const TF* v_fluxtop;
    TF visc;
};

// vertical advective flux of a cell-centred-in-z quantity (u, v, s) through face f:
// column values c[-3..+2] relative to the face (c[-1] below, c[0] above the face).
template <typename TF>
__device__ __forceinline__ TF vflux_face(const int order, const TF vel,
        const TF* __restrict__ q, const long long idx_above, const long long kk)
{
    // idx_above = index of the cell just above the face
    if (order == 6)
        return flux65(vel, q[idx_above - 3 * kk], q[idx_above - 2 * kk], q[idx_above - kk],
                           q[idx_above], q[idx_above + kk], q[idx_above + 2 * kk]);
    if (order == 4)
        return flux43(vel, q[idx_above - 2 * kk], q[idx_above - kk], q[idx_above], q[idx_above + kk]);
    if (order == 2)
        return flux2(vel, q[idx_above - kk], q[idx_above]);
    return TF(0);
}

template <typename TF, bool ADV, bool DIFF, bool SURFACE, bool BUOY>
__global__ void __launch_bounds__(256) tend_uvw_kernel(const MomArgs<TF> a, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const int ii = 1;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ij = i + j * jj;
    const long long ijk = ij + k * kk;
    const TF* __restrict__ u = a.u; const TF* __restrict__ v = a.v; const TF* __restrict__ w = a.w;
    const TF dxi = g.dxi, dyi = g.dyi;
    const int ks = g.kstart, ke = g.kend;

    TF tu = TF(0), tv = TF(0), tw = TF(0);

    if (ADV)
    {
        // ---- u ----
        {
            const TF fe = flux65(interp2(u[ijk], u[ijk + ii]), u[ijk - 2], u[ijk - 1], u[ijk], u[ijk + 1], u[ijk + 2], u[ijk + 3]);
            const TF fw = flux65(interp2(u[ijk - ii], u[ijk]), u[ijk - 3], u[ijk - 2], u[ijk - 1], u[ijk], u[ijk + 1], u[ijk + 2]);
            const TF fn = flux65(interp2(v[ijk - ii + jj], v[ijk + jj]), u[ijk - 2 * jj], u[ijk - jj], u[ijk], u[ijk + jj], u[ijk + 2 * jj], u[ijk + 3 * jj]);
            const TF fs = flux65(interp2(v[ijk - ii], v[ijk]), u[ijk - 3 * jj], u[ijk - 2 * jj], u[ijk - jj], u[ijk], u[ijk + jj], u[ijk + 2 * jj]);
            const TF ft = g.rhorefh[k + 1] * vflux_face<TF>(vorder(k + 1, ks, ke), interp2(w[ijk - ii + kk], w[ijk + kk]), u, ijk + kk, kk);
            const TF fb = g.rhorefh[k    ] * vflux_face<TF>(vorder(k,     ks, ke), interp2(w[ijk - ii     ], w[ijk     ]), u, ijk,      kk);
            tu += -(fe - fw) * dxi - (fn - fs) * dyi - (ft - fb) / g.rhoref[k] * g.dzi[k];
        }
        // ---- v ----
        {
            const TF fe = flux65(interp2(u[ijk + ii - jj], u[ijk + ii]), v[ijk - 2], v[ijk - 1], v[ijk], v[ijk + 1], v[ijk + 2], v[ijk + 3]);
            const TF fw = flux65(interp2(u[ijk - jj], u[ijk]), v[ijk - 3], v[ijk - 2], v[ijk - 1], v[ijk], v[ijk + 1], v[ijk + 2]);
            const TF fn = flux65(interp2(v[ijk], v[ijk + jj]), v[ijk - 2 * jj], v[ijk - jj], v[ijk], v[ijk + jj], v[ijk + 2 * jj], v[ijk + 3 * jj]);
            const TF fs = flux65(interp2(v[ijk - jj], v[ijk]), v[ijk - 3 * jj], v[ijk - 2 * jj], v[ijk - jj], v[ijk], v[ijk + jj], v[ijk + 2 * jj]);
            const TF ft = g.rhorefh[k + 1] * vflux_face<TF>(vorder(k + 1, ks, ke), interp2(w[ijk - jj + kk], w[ijk + kk]), v, ijk + kk, kk);
            const TF fb = g.rhorefh[k    ] * vflux_face<TF>(vorder(k,     ks, ke), interp2(w[ijk - jj     ], w[ijk     ]), v, ijk,      kk);
            tv += -(fe - fw) * dxi - (fn - fs) * dyi - (ft - fb) / g.rhoref[k] * g.dzi[k];
        }
        // ---- w (faces kstart+1 .. kend-1) ----
        if (k > ks)
        {
            const TF fe = flux65(interp2(u[ijk + ii - kk], u[ijk + ii]), w[ijk - 2], w[ijk - 1], w[ijk], w[ijk + 1], w[ijk + 2], w[ijk + 3]);
            const TF fw = flux65(interp2(u[ijk - kk], u[ijk]), w[ijk - 3], w[ijk - 2], w[ijk - 1], w[ijk], w[ijk + 1], w[ijk + 2]);
            const TF fn = flux65(interp2(v[ijk + jj - kk], v[ijk + jj]), w[ijk - 2 * jj], w[ijk - jj], w[ijk], w[ijk + jj], w[ijk + 2 * jj], w[ijk + 3 * jj]);
            const TF fs = flux65(interp2(v[ijk - kk], v[ijk]), w[ijk - 3 * jj], w[ijk - 2 * jj], w[ijk - jj], w[ijk], w[ijk + jj], w[ijk + 2 * jj]);
            // vertical fluxes live on cell centres c = k (top) and c = k-1 (bottom); the order is set by the
            // distance of the centre to the walls: centres kstart and kend-1 are 2nd order.
            const TF ft = g.rhoref[k    ] * vflux_face<TF>(vorder(k,     ks - 1, ke), interp2(w[ijk     ], w[ijk + kk]), w, ijk + kk, kk);
            const TF fb = g.rhoref[k - 1] * vflux_face<TF>(vorder(k - 1, ks - 1, ke), interp2(w[ijk - kk], w[ijk     ]), w, ijk,      kk);
            tw += -(fe - fw) * dxi - (fn - fs) * dyi - (ft - fb) / g.rhorefh[k] * g.dzhi[k];
        }
    }

    if (DIFF)
    {
        const TF* __restrict__ e = a.evisc;
        const TF visc = a.visc;
        const TF q = TF(0.25);
        const bool bot = SURFACE && (k == ks);
        const bool top = SURFACE && (k == ke - 1);
        // ---- u ----
        {
            const TF evisce = e[ijk] + visc;
            const TF eviscw = e[ijk - ii] + visc;
            const TF eviscn = q * (e[ijk - ii] + e[ijk] + e[ijk - ii + jj] + e[ijk + jj]) + visc;
            const TF eviscs = q * (e[ijk - ii - jj] + e[ijk - jj] + e[ijk - ii] + e[ijk]) + visc;
            TF d = (evisce * (u[ijk + ii] - u[ijk]) * dxi - eviscw * (u[ijk] - u[ijk - ii]) * dxi) * TF(2.) * dxi
                 + (eviscn * ((u[ijk + jj] - u[ijk]) * dyi + (v[ijk + jj] - v[ijk - ii + jj]) * dxi)
                  - eviscs * ((u[ijk] - u[ijk - jj]) * dyi + (v[ijk] - v[ijk - ii]) * dxi)) * dyi;
            TF ft, fb;
            if (top) ft = -g.rhorefh[ke] * a.u_fluxtop[ij];
            else
            {
                const TF evisct = q * (e[ijk - ii] + e[ijk] + e[ijk - ii + kk] + e[ijk + kk]) + visc;
                ft = g.rhorefh[k + 1] * evisct * ((u[ijk + kk] - u[ijk]) * g.dzhi[k + 1] + (w[ijk + kk] - w[ijk - ii + kk]) * dxi);
            }
            if (bot) fb = -g.rhorefh[ks] * a.u_fluxbot[ij];
            else
            {
                const TF eviscb = q * (e[ijk - ii - kk] + e[ijk - kk] + e[ijk - ii] + e[ijk]) + visc;
                fb = g.rhorefh[k] * eviscb * ((u[ijk] - u[ijk - kk]) * g.dzhi[k] + (w[ijk] - w[ijk - ii]) * dxi);
            }
            tu += d + (ft - fb) / g.rhoref[k] * g.dzi[k];
        }
        // ---- v ----
        {
            const TF evisce = q * (e[ijk - jj] + e[ijk] + e[ijk + ii - jj] + e[ijk + ii]) + visc;
            const TF eviscw = q * (e[ijk - ii - jj] + e[ijk - ii] + e[ijk - jj] + e[ijk]) + visc;
            const TF eviscn = e[ijk] + visc;
            const TF eviscs = e[ijk - jj] + visc;
            TF d = (evisce * ((v[ijk + ii] - v[ijk]) * dxi + (u[ijk + ii] - u[ijk + ii - jj]) * dyi)
                  - eviscw * ((v[ijk] - v[ijk - ii]) * dxi + (u[ijk] - u[ijk - jj]) * dyi)) * dxi
                 + (eviscn * (v[ijk + jj] - v[ijk]) * dyi - eviscs * (v[ijk] - v[ijk - jj]) * dyi) * TF(2.) * dyi;
            TF ft, fb;
            if (top) ft = -g.rhorefh[ke] * a.v_fluxtop[ij];
            else
            {
                const TF evisct = q * (e[ijk - jj] + e[ijk] + e[ijk + kk - jj] + e[ijk + kk]) + visc;
                ft = g.rhorefh[k + 1] * evisct * ((v[ijk + kk] - v[ijk]) * g.dzhi[k + 1] + (w[ijk + kk] - w[ijk - jj + kk]) * dyi);
            }
            if (bot) fb = -g.rhorefh[ks] * a.v_fluxbot[ij];
            else
            {
                const TF eviscb = q * (e[ijk - kk - jj] + e[ijk - kk] + e[ijk - jj] + e[ijk]) + visc;
                fb = g.rhorefh[k] * eviscb * ((v[ijk] - v[ijk - kk]) * g.dzhi[k] + (w[ijk] - w[ijk - jj]) * dyi);
            }
            tv += d + (ft - fb) / g.rhoref[k] * g.dzi[k];
        }
        // ---- w ----
        if (k > ks)
        {
            const TF evisce = q * (e[ijk - kk] + e[ijk] + e[ijk + ii - kk] + e[ijk + ii]) + visc;
            const TF eviscw = q * (e[ijk - ii - kk] + e[ijk - ii] + e[ijk - kk] + e[ijk]) + visc;
            const TF eviscn = q * (e[ijk - kk] + e[ijk] + e[ijk + jj - kk] + e[ijk + jj]) + visc;
            const TF eviscs = q * (e[ijk - jj - kk] + e[ijk - jj] + e[ijk - kk] + e[ijk]) + visc;
            const TF evisct = e[ijk] + visc;
            const TF eviscb = e[ijk - kk] + visc;
            const TF dzhi = g.dzhi[k];
            tw += (evisce * ((w[ijk + ii] - w[ijk]) * dxi + (u[ijk + ii] - u[ijk + ii - kk]) * dzhi)
                 - eviscw * ((w[ijk] - w[ijk - ii]) * dxi + (u[ijk] - u[ijk - kk]) * dzhi)) * dxi
                + (eviscn * ((w[ijk + jj] - w[ijk]) * dyi + (v[ijk + jj] - v[ijk + jj - kk]) * dzhi)
                 - eviscs * ((w[ijk] - w[ijk - jj]) * dyi + (v[ijk] - v[ijk - kk]) * dzhi)) * dyi
                + (g.rhoref[k] * evisct * (w[ijk + kk] - w[ijk]) * g.dzi[k]
                 - g.rhoref[k - 1] * eviscb * (w[ijk] - w[ijk - kk]) * g.dzi[k - 1]) / g.rhorefh[k] * TF(2.) * dzhi;
        }
    }

    if (BUOY)
    {
        if (k > ks)
            tw += TF(GRAV) / g.threfh[k] * (interp2(a.th[ijk - kk], a.th[ijk]) - g.threfh[k]);
    }

    a.ut[ijk] += tu;
    a.vt[ijk] += tv;
    if (k > ks) a.wt[ijk] += tw;
}

// ------------------------------------------------------------------------------------------
// Scalar tendency: Advec_2i5 advec_s + Diff_kernels::diff_c.
// ------------------------------------------------------------------------------------------
template <typename TF>
struct ScalArgs
{
    TF* st;
    const TF* s;
    const TF* u; const TF* v; const TF* w;
    const TF* evisc;
    const TF* fluxbot; const TF* fluxtop;
    TF visc, tPr;
    TF dxidxi, dyidyi;
};

template <typename TF, bool ADV, bool DIFF, bool SURFACE>
__global__ void __launch_bounds__(256) tend_s_kernel(const ScalArgs<TF> a, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const int ii = 1;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ij = i + j * jj;
    const long long ijk = ij + k * kk;
    const TF* __restrict__ s = a.s;
    const int ks = g.kstart, ke = g.kend;
    TF ts = TF(0);
    if (ADV)
    {
        const TF* __restrict__ u = a.u; const TF* __restrict__ v = a.v; const TF* __restrict__ w = a.w;
        const TF fe = flux65(u[ijk + ii], s[ijk - 2], s[ijk - 1], s[ijk], s[ijk + 1], s[ijk + 2], s[ijk + 3]);
        const TF fw = flux65(u[ijk], s[ijk - 3], s[ijk - 2], s[ijk - 1], s[ijk], s[ijk + 1], s[ijk + 2]);
        const TF fn = flux65(v[ijk + jj], s[ijk - 2 * jj], s[ijk - jj], s[ijk], s[ijk + jj], s[ijk + 2 * jj], s[ijk + 3 * jj]);
        const TF fs = flux65(v[ijk], s[ijk - 3 * jj], s[ijk - 2 * jj], s[ijk - jj], s[ijk], s[ijk + jj], s[ijk + 2 * jj]);
        const TF ft = g.rhorefh[k + 1] * vflux_face<TF>(vorder(k + 1, ks, ke), w[ijk + kk], s, ijk + kk, kk);
        const TF fb = g.rhorefh[k    ] * vflux_face<TF>(vorder(k,     ks, ke), w[ijk     ], s, ijk,      kk);
        ts += -(fe - fw) * g.dxi - (fn - fs) * g.dyi - (ft - fb) / g.rhoref[k] * g.dzi[k];
    }
    if (DIFF)
    {
        const TF* __restrict__ e = a.evisc;
        const TF h = TF(0.5);
        const TF tPr_i = TF(1) / a.tPr;
        const TF visc = a.visc;
        const bool bot = SURFACE && (k == ks);
        const bool top = SURFACE && (k == ke - 1);
        const TF evisce = h * (e[ijk] + e[ijk + ii]) * tPr_i + visc;
        const TF eviscw = h * (e[ijk - ii] + e[ijk]) * tPr_i + visc;
        const TF eviscn = h * (e[ijk] + e[ijk + jj]) * tPr_i + visc;
        const TF eviscs = h * (e[ijk - jj] + e[ijk]) * tPr_i + visc;
        TF d = (evisce * (s[ijk + ii] - s[ijk]) - eviscw * (s[ijk] - s[ijk - ii])) * a.dxidxi
             + (eviscn * (s[ijk + jj] - s[ijk]) - eviscs * (s[ijk] - s[ijk - jj])) * a.dyidyi;
        TF ft, fb;
        if (top) ft = -g.rhorefh[ke] * a.fluxtop[ij];
        else
        {
            const TF evisct = h * (e[ijk] + e[ijk + kk]) * tPr_i + visc;
            ft = g.rhorefh[k + 1] * evisct * (s[ijk + kk] - s[ijk]) * g.dzhi[k + 1];
        }
        if (bot) fb = -g.rhorefh[ks] * a.fluxbot[ij];
        else
        {
            const TF eviscb = h * (e[ijk - kk] + e[ijk]) * tPr_i + visc;
            fb = g.rhorefh[k] * eviscb * (s[ijk] - s[ijk - kk]) * g.dzhi[k];
        }
        ts += d + (ft - fb) / g.rhoref[k] * g.dzi[k];
    }
    a.st[ijk] += ts;
}

// ------------------------------------------------------------------------------------------
// Flux-limited scalar advection (Koren 1993; reference include/advec_monotonic.h:28-202), used by Advec_2i5 for the
// scalars of `fluxlimit_list` (src/advec_2i5.cxx:1046-1056).  VARIANT 0: flux_lim, 1: flux_lim_bot (first-order upwind
// from below), 2: flux_lim_top (first-order upwind from above).  Faces kstart and kend carry no flux.
// ------------------------------------------------------------------------------------------
template <typename TF> __device__ __forceinline__ TF lim_eps();
template <> __device__ __forceinline__ double lim_eps<double>() { return 2.220446049250313e-16; }
template <> __device__ __forceinline__ float lim_eps<float>() { return 1.1920929e-07f; }
template <typename TF> __device__ __forceinline__ TF maxf_(TF a, TF b) { return a > b ? a : b; }
template <typename TF> __device__ __forceinline__ TF minf_(TF a, TF b) { return a < b ? a : b; }

template <typename TF>
__device__ __forceinline__ TF lim_branch(const TF vel, const TF a2, const TF a1, const TF b1)
{
    // vel*(a1 + phi/2 (a1 - a2)), phi = max(0, min(two_r, (1 + two_r)/3, 2)), two_r = 2 (b1 - a1) / guarded(a1 - a2)
    const TF d = a1 - a2;
    const TF mag = maxf_(absf(d), lim_eps<TF>());
    const TF denom = (d < TF(0) || (d == TF(0) && signbit(d))) ? -mag : mag;          // copysign(1, d) * max(|d|, eps)
    const TF two_r = TF(2.) * (b1 - a1) / denom;
    const TF phi = maxf_(TF(0.), minf_(two_r, minf_(TF(1. / 3.) * (TF(1.) + two_r), TF(2.))));
    return vel * (a1 + TF(0.5) * phi * (a1 - a2));
}

template <typename TF, int VARIANT>
__device__ __forceinline__ TF flux_lim(const TF vel, const TF sm2, const TF sm1, const TF sp1, const TF sp2)
{
    if (vel >= TF(0.)) return VARIANT == 1 ? vel * sm1 : lim_branch<TF>(vel, sm2, sm1, sp1);
    return VARIANT == 2 ? vel * sp1 : lim_branch<TF>(vel, sp2, sp1, sm1);
}

template <typename TF>
__global__ void __launch_bounds__(256) advec_s_lim_kernel(TF* __restrict__ st, const TF* __restrict__ s,
        const TF* __restrict__ u, const TF* __restrict__ v, const TF* __restrict__ w, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ijk = i + j * jj + k * kk;
    const int ks = g.kstart, ke = g.kend;
    const TF hx = (flux_lim<TF, 0>(u[ijk + 1], s[ijk - 1], s[ijk], s[ijk + 1], s[ijk + 2])
                 - flux_lim<TF, 0>(u[ijk], s[ijk - 2], s[ijk - 1], s[ijk], s[ijk + 1])) * g.dxi;
    const TF hy = (flux_lim<TF, 0>(v[ijk + jj], s[ijk - jj], s[ijk], s[ijk + jj], s[ijk + 2 * jj])
                 - flux_lim<TF, 0>(v[ijk], s[ijk - 2 * jj], s[ijk - jj], s[ijk], s[ijk + jj])) * g.dyi;
    // vertical face fluxes: face f lies below level f; variants next to the walls
    auto face = [&](const int f, const long long o) -> TF {       // o = index of the cell above the face
        const TF wv = w[o];
        if (f == ks + 1) return flux_lim<TF, 1>(wv, s[o - 2 * kk], s[o - kk], s[o], s[o + kk]);
        if (f == ke - 1) return flux_lim<TF, 2>(wv, s[o - 2 * kk], s[o - kk], s[o], s[o + kk]);
        return flux_lim<TF, 0>(wv, s[o - 2 * kk], s[o - kk], s[o], s[o + kk]);
    };
    TF vert;
    if (k == ks) vert = g.rhorefh[k + 1] * face(k + 1, ijk + kk);
    else if (k == ke - 1) vert = -(g.rhorefh[k] * face(k, ijk));
    else vert = g.rhorefh[k + 1] * face(k + 1, ijk + kk) - g.rhorefh[k] * face(k, ijk);
    st[ijk] += -hx - hy - vert / g.rhoref[k] * g.dzi[k];
}

// thermo_dry buoyancy tendency on w (src/thermo_dry.cxx:165-179)
template <typename TF>
__global__ void buoyancy_kernel(TF* __restrict__ wt, const TF* __restrict__ th, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + 1 + blockIdx.z;
    if (i >= g.iend || j >= g.jend || k >= g.kend) return;
    const long long ijk = i + (long long)j * g.icells + k * g.ijcells;
    wt[ijk] += TF(GRAV) / g.threfh[k] * (interp2(th[ijk - g.ijcells], th[ijk]) - g.threfh[k]);
}

// calc_N2 as a stand-alone field (Thermo_dry::get_thermo_field("N2"))
template <typename TF>
__global__ void n2_kernel(TF* __restrict__ n2, const TF* __restrict__ th, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long ijk = i + (long long)j * g.icells + k * g.ijcells;
    n2[ijk] = TF(GRAV) / g.thref[k] * TF(0.5) * (th[ijk + g.ijcells] - th[ijk - g.ijcells]) * g.dzi[k];
}

// ------------------------------------------------------------------------------------------
// Reductions: CFL (Advec_2i5 calc_cfl), diffusion number (calc_dnmul), divergence (Pres_2).
// mode 0 = cfl, 1 = dnmul, 2 = divergence.  Result = max over interior, atomically into *out.
// ------------------------------------------------------------------------------------------
template <typename TF, int MODE>
__global__ void __launch_bounds__(256) reduce_kernel(const TF* __restrict__ u, const TF* __restrict__ v, const TF* __restrict__ w,
        const GridDev<TF> g, const TF p0, const TF p1, const TF p2, double* __restrict__ out)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    TF val = TF(0);
    if (i < g.iend && j < g.jend)
    {
        const long long jj = g.icells, kk = g.ijcells;
        const long long ijk = i + j * jj + k * kk;
        if (MODE == 0)
        {
            const int o = vorder(k, g.kstart - 1, g.kend);   // centre-based order (2,4,6)
            TF wi;
            if (o == 2) wi = interp2(w[ijk], w[ijk + kk]);
            else if (o == 4) wi = interp4_ws(w[ijk - kk], w[ijk], w[ijk + kk], w[ijk + 2 * kk]);
            else wi = interp6_ws(w[ijk - 2 * kk], w[ijk - kk], w[ijk], w[ijk + kk], w[ijk + 2 * kk], w[ijk + 3 * kk]);
            val = absf(interp6_ws(u[ijk - 2], u[ijk - 1], u[ijk], u[ijk + 1], u[ijk + 2], u[ijk + 3])) * g.dxi
                + absf(interp6_ws(v[ijk - 2 * jj], v[ijk - jj], v[ijk], v[ijk + jj], v[ijk + 2 * jj], v[ijk + 3 * jj])) * g.dyi
                + absf(wi) * g.dzi[k];
        }
        else if (MODE == 1)
        {
            // u = evisc; p0 = 1/min(1,tPr); p1 = 1/dx^2; p2 = 1/dy^2
            val = absf(u[ijk] * p0 * (p1 + p2 + g.dzi[k] * g.dzi[k]));
        }
        else
        {
            val = absf(g.rhoref[k] * ((u[ijk + 1] - u[ijk]) * g.dxi + (v[ijk + jj] - v[ijk]) * g.dyi)
                     + (g.rhorefh[k + 1] * w[ijk + kk] - g.rhorefh[k] * w[ijk]) * g.dzi[k]);
        }
    }
    block_max_to_global<TF>(val, out);
}

// ------------------------------------------------------------------------------------------
// Pressure: rhs (Pres_2::input) into a compact array with row pitch `pitch`; tendency
// correction (Pres_2::output); RK3 update (Timeloop rk3).
// ------------------------------------------------------------------------------------------
template <typename TF>
struct PresArgs
{
    TF* ut; TF* vt; TF* wt;
    const TF* u; const TF* v; const TF* w;
    TF* p;
};

// The periodic wrap of ut (east) and vt (north) is taken by index instead of a halo fill of the
// tendencies (the reference does boundary_cyclic.exec(ut, East_west) / (vt, North_south) first;
// single-GPU only -- the multi-GPU path exchanges those halos).
template <typename TF>
__global__ void __launch_bounds__(256) pres_in_kernel(const PresArgs<TF> a, TF* __restrict__ rhs, const long long pitch_j,
        const long long pitch_k, const TF dti, const GridDev<TF> g, const int wrap)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ijk = i + j * jj + k * kk;
    const long long ie = (wrap && i + 1 == g.iend) ? ijk + 1 - g.imax : ijk + 1;
    const long long jn = (wrap && j + 1 == g.jend) ? ijk + (1 - g.jmax) * jj : ijk + jj;
    const TF val = g.rhoref[k] * ((a.ut[ie] + a.u[ie] * dti) - (a.ut[ijk] + a.u[ijk] * dti)) * g.dxi
                 + g.rhoref[k] * ((a.vt[jn] + a.v[jn] * dti) - (a.vt[ijk] + a.v[ijk] * dti)) * g.dyi
                 + (g.rhorefh[k + 1] * (a.wt[ijk + kk] + a.w[ijk + kk] * dti)
                  - g.rhorefh[k    ] * (a.wt[ijk     ] + a.w[ijk     ] * dti)) * g.dzi[k];
    rhs[(i - g.istart) + (j - g.jstart) * pitch_j + (k - g.kstart) * pitch_k] = val;
}

template <typename TF>
__global__ void __launch_bounds__(256) pres_out_kernel(const PresArgs<TF> a, const GridDev<TF> g)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + blockIdx.z;
    if (i >= g.iend || j >= g.jend) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ijk = i + j * jj + k * kk;
    const TF pc = a.p[ijk];
    a.ut[ijk] -= (pc - a.p[ijk - 1]) * g.dxi;
    a.vt[ijk] -= (pc - a.p[ijk - jj]) * g.dyi;
    a.wt[ijk] -= (pc - a.p[ijk - kk]) * g.dzhi[k];
}

// rk3: a += cB*dt*at over the interior; at *= cA[next] over the interior, or at = 0 over ALL cells
// (ghosts too, as the CPU reference does) when the sub-step wraps.  One thread per cell (all cells).
template <typename TF>
__global__ void __launch_bounds__(256) rk3_kernel(TF* __restrict__ a, TF* __restrict__ at, const TF cbdt, const TF ca_next,
        const int wrap, const GridDev<TF> g)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= g.icells || j >= g.jcells) return;
    const long long ijk = i + (long long)j * g.icells + k * g.ijcells;
    const bool interior = i >= g.istart && i < g.iend && j >= g.jstart && j < g.jend && k >= g.kstart && k < g.kend;
    if (interior)
    {
        const TF t = at[ijk];
        a[ijk] += cbdt * t;
        at[ijk] = wrap ? TF(0) : t * ca_next;
    }
    else if (wrap)
        at[ijk] = TF(0);
}

// Fused pressure correction + RK3 for the three momentum components: reads p once,
// RMW u,v,w and ut,vt,wt once (SURVEY section 8d stage v).
template <typename TF>
__global__ void __launch_bounds__(256) pres_out_rk3_kernel(const PresArgs<TF> a, TF* __restrict__ u, TF* __restrict__ v, TF* __restrict__ w,
        const TF cbdt, const TF ca_next, const int wrap, const GridDev<TF> g)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= g.icells || j >= g.jcells) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ijk = i + j * jj + k * kk;
    const bool interior = i >= g.istart && i < g.iend && j >= g.jstart && j < g.jend && k >= g.kstart && k < g.kend;
    if (interior)
    {
        const TF pc = a.p[ijk];
        const TF tu = a.ut[ijk] - (pc - a.p[ijk - 1]) * g.dxi;
        const TF tv = a.vt[ijk] - (pc - a.p[ijk - jj]) * g.dyi;
        const TF tw = a.wt[ijk] - (pc - a.p[ijk - kk]) * g.dzhi[k];
        u[ijk] += cbdt * tu;
        v[ijk] += cbdt * tv;
        w[ijk] += cbdt * tw;
        a.ut[ijk] = wrap ? TF(0) : tu * ca_next;
        a.vt[ijk] = wrap ? TF(0) : tv * ca_next;
        a.wt[ijk] = wrap ? TF(0) : tw * ca_next;
    }
    else if (wrap)
    {
        a.ut[ijk] = TF(0); a.vt[ijk] = TF(0); a.wt[ijk] = TF(0);
    }
}

// ---- two cells per thread (one 2*sizeof(TF) vector access per array): same arithmetic as rk3_kernel / pres_out_rk3_kernel.
// Needs an even icells and 2*sizeof(TF)-aligned arrays (every row then starts on a vector boundary).  In single precision a
// thread of the scalar kernels has only 4 bytes in flight per array, which is what kept them at 58 % / 74 % of the HBM rate.
template <typename TF>
__global__ void __launch_bounds__(256) rk3_v2_kernel(TF* __restrict__ a, TF* __restrict__ at, const TF cbdt, const TF ca_next,
        const int wrap, const GridDev<TF> g)
{
    typedef typename V2T<TF>::type V2;
    const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= g.icells || j >= g.jcells) return;
    const long long ijk = i + (long long)j * g.icells + k * g.ijcells;
    const bool row = j >= g.jstart && j < g.jend && k >= g.kstart && k < g.kend;
    const bool in0 = row && i >= g.istart && i < g.iend, in1 = row && i + 1 >= g.istart && i + 1 < g.iend;
    if (!in0 && !in1 && !wrap) return;
    V2 t = *reinterpret_cast<const V2*>(at + ijk);
    if (in0 || in1)
    {
        V2 x = *reinterpret_cast<const V2*>(a + ijk);
        if (in0) x.x += cbdt * t.x;
        if (in1) x.y += cbdt * t.y;
        *reinterpret_cast<V2*>(a + ijk) = x;
    }
    t.x = wrap ? TF(0) : (in0 ? t.x * ca_next : t.x);
    t.y = wrap ? TF(0) : (in1 ? t.y * ca_next : t.y);
    *reinterpret_cast<V2*>(at + ijk) = t;
}

template <typename TF>
__global__ void __launch_bounds__(256) pres_out_rk3_v2_kernel(const PresArgs<TF> a, TF* __restrict__ u, TF* __restrict__ v, TF* __restrict__ w,
        const TF cbdt, const TF ca_next, const int wrap, const GridDev<TF> g)
{
    typedef typename V2T<TF>::type V2;
    const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= g.icells || j >= g.jcells) return;
    const long long jj = g.icells, kk = g.ijcells;
    const long long ijk = i + j * jj + k * kk;
    const bool row = j >= g.jstart && j < g.jend && k >= g.kstart && k < g.kend;
    const bool in0 = row && i >= g.istart && i < g.iend, in1 = row && i + 1 >= g.istart && i + 1 < g.iend;
    auto LD = [](const TF* p) -> V2 { return *reinterpret_cast<const V2*>(p); };
    auto ST = [](TF* p, const V2 x) { *reinterpret_cast<V2*>(p) = x; };
    if (in0 || in1)
    {
        const V2 pc = LD(a.p + ijk), ps = LD(a.p + ijk - jj), pb = LD(a.p + ijk - kk);
        const TF pw = a.p[ijk - 1];                 // in0 implies i >= istart >= 1; with !in0 the value is not used
        V2 tu = LD(a.ut + ijk), tv = LD(a.vt + ijk), tw = LD(a.wt + ijk);
        V2 xu = LD(u + ijk), xv = LD(v + ijk), xw = LD(w + ijk);
        const TF dzhi = g.dzhi[k];
        if (in0)
        {
            const TF nu = tu.x - (pc.x - pw) * g.dxi, nv = tv.x - (pc.x - ps.x) * g.dyi, nw = tw.x - (pc.x - pb.x) * dzhi;
            xu.x += cbdt * nu; xv.x += cbdt * nv; xw.x += cbdt * nw;
            tu.x = nu * ca_next; tv.x = nv * ca_next; tw.x = nw * ca_next;
        }
        if (in1)
        {
            const TF nu = tu.y - (pc.y - pc.x) * g.dxi, nv = tv.y - (pc.y - ps.y) * g.dyi, nw = tw.y - (pc.y - pb.y) * dzhi;
            xu.y += cbdt * nu; xv.y += cbdt * nv; xw.y += cbdt * nw;
            tu.y = nu * ca_next; tv.y = nv * ca_next; tw.y = nw * ca_next;
        }
        ST(u + ijk, xu); ST(v + ijk, xv); ST(w + ijk, xw);
        if (wrap) { tu.x = tu.y = tv.x = tv.y = tw.x = tw.y = TF(0); }
        ST(a.ut + ijk, tu); ST(a.vt + ijk, tv); ST(a.wt + ijk, tw);
    }
    else if (wrap)
    {
        V2 z; z.x = TF(0); z.y = TF(0);
        ST(a.ut + ijk, z); ST(a.vt + ijk, z); ST(a.wt + ijk, z);
    }
}

} // namespace mhh
