// mhhb200 -- Thermo_moist<TF> (liquid-water potential temperature thl + total water qt; reference src/thermo_moist.cxx and
// include/thermo_moist_functions.h).
//   saturation adjustment sat_adjust            functions.h:164-268   (Newton on T; liquid above T0, mixed phase below)
//   esat / qsat / dqsatdT (liquid: 10th-order Taylor polynomial of Arden Buck; ice: exponential)   functions.h:75-149
//   calc_base_state                             functions.h:271-340   hydrostatic pressure, exner, thv, rho at full / half levels
//   calc_buoyancy_tend_2nd                      src/thermo_moist.cxx:77-120
//   calc_buoyancy / calc_liquid_water / calc_N2 src/thermo_moist.cxx:122-168, 230-250, 459-475   (get_thermo_field b / ql / N2)
//   calc_buoyancy_bot / calc_buoyancy_fluxbot   src/thermo_moist.cxx:637-693
//   Field3d_operators::calc_mean_profile        src/field3d_operators.cxx:45-66
// Where the reference's GPU build copies the mean profiles to the host, integrates the base state there and copies eight
// profiles back on EVERY sub-step (src/thermo_moist.cu:914-943: two blocking D2H, one H2D), this path keeps the update on the
// device: one reduction kernel for the two mean profiles and one single-lane kernel for the (inherently serial, kmax-long)
// hydrostatic integration, both stream-ordered -- no host round trip, and the sub-step stays capturable in a CUDA graph.
// The buoyancy kernel is HBM-bound: thl and qt read once (the level below comes out of L2), wt read and written once = 4
// array passes; the exner function of the level is evaluated by one lane per CTA.
#pragma once
#include "common.cuh"

namespace mhh {

// The arithmetic below is __host__ __device__: tests/test_moist_hostcheck.py compiles these very functions for the CPU and holds
// them bit for bit against the oracle (the kernels around them only index).
#define MHH_HD __host__ __device__ __forceinline__

// Constants::* (include/constants.h:29-39, 73-84), formed in TF like the reference's `template<typename TF> constexpr TF`
template <typename TF>
struct MoistC
{
    static constexpr TF grav = TF(9.81), Rd = TF(287.04), Rv = TF(461.5), cp = TF(1005), Lv = TF(2.501e6), Lf = TF(3.337e5);
    static constexpr TF Ls = Lv + Lf, T0 = TF(273.15), p0 = TF(1.e5), ep = Rd / Rv;
};

MHH_HD double m_pow(double a, double b) { return pow(a, b); }
MHH_HD float  m_pow(float a, float b) { return powf(a, b); }
MHH_HD double m_abs(double a) { return fabs(a); }
MHH_HD float  m_abs(float a) { return fabsf(a); }
MHH_HD double m_exp(double a) { return exp(a); }
MHH_HD float  m_exp(float a) { return expf(a); }
MHH_HD double m_min(double a, double b) { return fmin(a, b); }
MHH_HD float  m_min(float a, float b) { return fminf(a, b); }
MHH_HD double m_max(double a, double b) { return fmax(a, b); }
MHH_HD float  m_max(float a, float b) { return fmaxf(a, b); }

template <typename TF> MHH_HD TF moist_exner(TF p)
{ return m_pow(p / MoistC<TF>::p0, MoistC<TF>::Rd / MoistC<TF>::cp); }

template <typename TF> MHH_HD TF moist_esat_liq(TF T)
{
    const TF x = m_min(m_max(TF(-75.), T - MoistC<TF>::T0), TF(50.));
    return TF(+6.1121000000E+02) + x * (TF(+4.4393067270E+01) + x * (TF(+1.4279398448E+00) + x * (TF(+2.6415206946E-02)
         + x * (TF(+3.0291749160E-04) + x * (TF(+2.1159987257E-06) + x * (TF(+7.5015702516E-09) + x * (TF(-1.5604873363E-12)
         + x * (TF(-9.9726710231E-14) + x * (TF(-4.8165754883E-17) + x * TF(+1.3839187032E-18))))))))));
}
template <typename TF> MHH_HD TF moist_esat_ice(TF T)
{
    const TF x = m_min(m_max(TF(-100.), T - MoistC<TF>::T0), TF(50.));
    return TF(611.15) * m_exp(TF(22.452) * x / (TF(272.55) + x));
}
template <typename TF> MHH_HD TF moist_qsat_from_es(TF p, TF es)
{ return MoistC<TF>::ep * es / (p - (TF(1.) - MoistC<TF>::ep) * es); }
template <typename TF> MHH_HD TF moist_water_fraction(TF T)
{ return m_max(TF(0.), m_min((T - TF(233.15)) / (MoistC<TF>::T0 - TF(233.15)), TF(1.))); }
// dqsat/dT by Clausius-Clapeyron (functions.h:121-141); L = Lv (liquid) or Ls (ice)
template <typename TF> MHH_HD TF moist_dqsatdT_from_es(TF p, TF T, TF es, TF L)
{
    typedef MoistC<TF> C;
    const TF den = p - es * (TF(1.) - C::ep);
    return (C::ep / den + (TF(1.) - C::ep) * C::ep * es / (den * den)) * L * es / (C::Rv * (T * T));
}
template <typename TF> MHH_HD TF moist_virtual_temperature(TF exn, TF thl, TF qt, TF ql, TF qi)
{
    typedef MoistC<TF> C;
    const TF th = thl + C::Lv * ql / (C::cp * exn) + C::Ls * qi / (C::cp * exn);
    return th * (TF(1.) - (TF(1.) - C::Rv / C::Rd) * qt - C::Rv / C::Rd * (ql + qi));
}
template <typename TF> MHH_HD TF moist_buoyancy(TF exn, TF thl, TF qt, TF ql, TF qi, TF thvref)
{ return MoistC<TF>::grav * (moist_virtual_temperature(exn, thl, qt, ql, qi) - thvref) / thvref; }
template <typename TF> MHH_HD TF moist_buoyancy_no_ql(TF thl, TF qt, TF thvref)
{
    typedef MoistC<TF> C;
    return C::grav * (thl * (TF(1.) - (TF(1.) - C::Rv / C::Rd) * qt) - thvref) / thvref;
}
template <typename TF> MHH_HD TF moist_buoyancy_flux_no_ql(TF thl, TF thlflux, TF qt, TF qtflux, TF thvref)
{
    typedef MoistC<TF> C;
    return C::grav / thvref * (thlflux * (TF(1.) - (TF(1.) - C::Rv / C::Rd) * qt) - (TF(1.) - C::Rv / C::Rd) * thl * qtflux);
}

template <typename TF> struct SatAdjust { TF ql, qi, t, qs; bool converged; };

// sat_adjust (functions.h:164-268).  `converged` is false where the reference throws ("Non-converging saturation adjustment").
template <typename TF>
MHH_HD SatAdjust<TF> moist_sat_adjust(TF thl, TF qt, TF p, TF exn)
{
    typedef MoistC<TF> C;
    const TF tl = thl * exn;
    TF qs = moist_qsat_from_es(p, moist_esat_liq(tl));
    SatAdjust<TF> ans{TF(0.), TF(0.), tl, qs, true};
    if (qt - qs <= TF(0.)) return ans;
    int niter = 0;
    const int nitermax = 10;
    TF tnr_old = TF(1.e9), tnr = tl;
    if (tl >= C::T0)
    {
        while (m_abs(tnr - tnr_old) / tnr_old > TF(1.e-5) && niter < nitermax)
        {
            ++niter;
            tnr_old = tnr;
            const TF es = moist_esat_liq(tnr);
            qs = moist_qsat_from_es(p, es);
            const TF f = tnr - tl - C::Lv / C::cp * (qt - qs);
            const TF f_prime = TF(1.) + C::Lv / C::cp * moist_dqsatdT_from_es(p, tnr, es, C::Lv);
            tnr -= f / f_prime;
        }
        qs = moist_qsat_from_es(p, moist_esat_liq(tnr));
        ans.ql = m_max(TF(0.), qt - qs);
    }
    else
    {
        while (m_abs(tnr - tnr_old) / tnr_old > TF(1.e-5) && niter < nitermax)
        {
            ++niter;
            tnr_old = tnr;
            const TF esl = moist_esat_liq(tnr), esi = moist_esat_ice(tnr);
            const TF alpha_w = moist_water_fraction(tnr), alpha_i = TF(1.) - alpha_w;
            qs = alpha_w * moist_qsat_from_es(p, esl) + alpha_i * moist_qsat_from_es(p, esi);
            const TF dalphadT = (alpha_w > TF(0.) && alpha_w < TF(1.)) ? TF(0.025) : TF(0.);
            const TF dqsatdT_w = moist_dqsatdT_from_es(p, tnr, esl, C::Lv);
            const TF dqsatdT_i = moist_dqsatdT_from_es(p, tnr, esi, C::Ls);
            const TF f = tnr - tl - alpha_w * C::Lv / C::cp * qt - alpha_i * C::Ls / C::cp * qt
                                  + alpha_w * C::Lv / C::cp * qs + alpha_i * C::Ls / C::cp * qs;
            const TF f_prime = TF(1.)
                - dalphadT * C::Lv / C::cp * qt + dalphadT * C::Ls / C::cp * qt
                + dalphadT * C::Lv / C::cp * qs - dalphadT * C::Ls / C::cp * qs
                + alpha_w * C::Lv / C::cp * dqsatdT_w
                + alpha_i * C::Ls / C::cp * dqsatdT_i;
            tnr -= f / f_prime;
        }
        const TF alpha_w = moist_water_fraction(tnr), alpha_i = TF(1.) - alpha_w;
        qs = alpha_w * moist_qsat_from_es(p, moist_esat_liq(tnr)) + alpha_i * moist_qsat_from_es(p, moist_esat_ice(tnr));
        const TF qlqi = m_max(TF(0.), qt - qs);
        ans.ql = alpha_w * qlqi;
        ans.qi = alpha_i * qlqi;
    }
    ans.t = tnr;
    ans.qs = qs;
    ans.converged = niter != nitermax;
    return ans;
}

// The eight base-state profiles of Thermo_moist's `bs` (kcells entries each); thvref / thvrefh alias the context's
// thref / threfh (the closures' N2 and the surface model read them from there).
template <typename TF>
struct MoistProfiles
{
    TF *pref, *prefh, *rho, *rhoh, *thv, *thvh, *ex, *exh;
};

// Field3d_operators::calc_mean_profile for two fields at once: one CTA per (level, field), double accumulation like the
// reference (its sum is sequential, this one a fixed-shape tree: deterministic, equal to rounding of the double sum).
template <typename TF>
__global__ void __launch_bounds__(256) moist_mean_profile_kernel(const TF* __restrict__ f0, const TF* __restrict__ f1,
        TF* __restrict__ m0, TF* __restrict__ m1, const GridDev<TF> g, const double n)
{
    const int k = blockIdx.x;
    const TF* __restrict__ f = blockIdx.y == 0 ? f0 : f1;
    TF* __restrict__ m = blockIdx.y == 0 ? m0 : m1;
    const long long base = (long long)k * g.ijcells;
    double s = 0.;
    const int nij = g.imax * g.jmax;
    for (int idx = threadIdx.x; idx < nij; idx += blockDim.x)
    {
        const int j = idx / g.imax, i = idx - j * g.imax;
        s += (double)f[base + (long long)(j + g.jstart) * g.icells + (i + g.istart)];
    }
    __shared__ double red[8];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        double t = 0.;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        m[k] = (TF)(t / n);
    }
}

// calc_base_state (functions.h:271-340): serial in k.  Returns the number of non-converged saturation adjustments.
template <typename TF>
MHH_HD int moist_base_state_serial(const MoistProfiles<TF> b, const TF* __restrict__ thlmean, const TF* __restrict__ qtmean,
        const TF pbot, const int kstart, const int kend, const TF* __restrict__ z, const TF* __restrict__ dz, const TF* __restrict__ dzh)
{
    typedef MoistC<TF> C;
    int bad = 0;
    const TF thlsurf = TF(0.5) * (thlmean[kstart - 1] + thlmean[kstart]);
    const TF qtsurf  = TF(0.5) * (qtmean[kstart - 1] + qtmean[kstart]);
    b.prefh[kstart] = pbot;
    b.exh[kstart] = moist_exner(pbot);
    SatAdjust<TF> ssa = moist_sat_adjust(thlsurf, qtsurf, b.prefh[kstart], b.exh[kstart]);
    bad += ssa.converged ? 0 : 1;
    b.thvh[kstart] = moist_virtual_temperature(b.exh[kstart], thlsurf, qtsurf, ssa.ql, ssa.qi);
    b.rhoh[kstart] = pbot / (C::Rd * b.exh[kstart] * b.thvh[kstart]);
    b.pref[kstart] = b.prefh[kstart] * m_exp(-C::grav * z[kstart] / (C::Rd * b.exh[kstart] * b.thvh[kstart]));
    for (int k = kstart + 1; k < kend + 1; ++k)
    {
        b.ex[k - 1] = moist_exner(b.pref[k - 1]);
        ssa = moist_sat_adjust(thlmean[k - 1], qtmean[k - 1], b.pref[k - 1], b.ex[k - 1]);
        bad += ssa.converged ? 0 : 1;
        b.thv[k - 1] = moist_virtual_temperature(b.ex[k - 1], thlmean[k - 1], qtmean[k - 1], ssa.ql, ssa.qi);
        b.rho[k - 1] = b.pref[k - 1] / (C::Rd * b.ex[k - 1] * b.thv[k - 1]);
        b.prefh[k] = b.prefh[k - 1] * m_exp(-C::grav * dz[k - 1] / (C::Rd * b.ex[k - 1] * b.thv[k - 1]));
        b.exh[k] = moist_exner(b.prefh[k]);
        const TF thli = TF(0.5) * (thlmean[k - 1] + thlmean[k]);
        const TF qti  = TF(0.5) * (qtmean[k - 1] + qtmean[k]);
        ssa = moist_sat_adjust(thli, qti, b.prefh[k], b.exh[k]);
        bad += ssa.converged ? 0 : 1;
        b.thvh[k] = moist_virtual_temperature(b.exh[k], thli, qti, ssa.ql, ssa.qi);
        b.rhoh[k] = b.prefh[k] / (C::Rd * b.exh[k] * b.thvh[k]);
        b.pref[k] = b.pref[k - 1] * m_exp(-C::grav * dzh[k] / (C::Rd * b.exh[k] * b.thvh[k]));
    }
    b.pref[kstart - 1] = TF(2.) * b.prefh[kstart] - b.pref[kstart];
    return bad;
}

// one lane: the integration is a kmax-long dependent chain
template <typename TF>
__global__ void moist_base_state_kernel(const MoistProfiles<TF> b, const TF* __restrict__ thlmean, const TF* __restrict__ qtmean,
        const TF pbot, const GridDev<TF> g, int* __restrict__ nonconv)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int bad = moist_base_state_serial<TF>(b, thlmean, qtmean, pbot, g.kstart, g.kend, g.z, g.dz, g.dzh);
    if (bad) atomicAdd(nonconv, bad);
}

// calc_buoyancy_tend_2nd (src/thermo_moist.cxx:77-120): wt += buoyancy of (thl, qt) interpolated to the half level, with the
// condensate of the saturation adjustment at that level's pressure.  One level per blockIdx.z (k = kstart+1 .. kend-1).
template <typename TF>
__global__ void __launch_bounds__(256) moist_buoyancy_tend_kernel(TF* __restrict__ wt, const TF* __restrict__ thl, const TF* __restrict__ qt,
        const TF* __restrict__ ph, const TF* __restrict__ thvrefh, const GridDev<TF> g, int* __restrict__ nonconv)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = g.kstart + 1 + blockIdx.z;
    __shared__ TF s_exnh;
    if (threadIdx.x == 0 && threadIdx.y == 0) s_exnh = moist_exner(ph[k]);
    __syncthreads();
    if (i >= g.iend || j >= g.jend) return;
    const TF exnh = s_exnh, p = ph[k];
    const long long ijk = i + (long long)j * g.icells + k * g.ijcells;
    const TF thlh = interp2(thl[ijk - g.ijcells], thl[ijk]);
    const TF qth  = interp2(qt[ijk - g.ijcells], qt[ijk]);
    const SatAdjust<TF> ssa = moist_sat_adjust(thlh, qth, p, exnh);
    if (!ssa.converged) atomicAdd(nonconv, 1);
    wt[ijk] += moist_buoyancy(exnh, thlh, qth, ssa.ql, ssa.qi, thvrefh[k]);
}

// get_thermo_field: MODE 0 = "b" (calc_buoyancy: every level, no condensate outside kstart..kend-1), 1 = "ql"
// (calc_liquid_water), 2 = "N2" (calc_N2); interior columns only, like the reference's loops.
template <typename TF, int MODE>
__global__ void __launch_bounds__(256) moist_field_kernel(TF* __restrict__ out, const TF* __restrict__ thl, const TF* __restrict__ qt,
        const TF* __restrict__ p, const TF* __restrict__ thvref, const GridDev<TF> g, int* __restrict__ nonconv)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = MODE == 0 ? (int)blockIdx.z : g.kstart + (int)blockIdx.z;
    __shared__ TF s_ex;
    if (MODE != 2)
    {
        if (threadIdx.x == 0 && threadIdx.y == 0) s_ex = moist_exner(p[k]);
        __syncthreads();
    }
    if (i >= g.iend || j >= g.jend) return;
    const long long ijk = i + (long long)j * g.icells + k * g.ijcells;
    if (MODE == 2)
    {
        out[ijk] = MoistC<TF>::grav / thvref[k] * TF(0.5) * (thl[ijk + g.ijcells] - thl[ijk - g.ijcells]) * g.dzi[k];
        return;
    }
    const TF ex = s_ex;
    TF ql = TF(0.), qi = TF(0.);
    if (MODE == 1 || (k >= g.kstart && k < g.kend))
    {
        const SatAdjust<TF> ssa = moist_sat_adjust(thl[ijk], qt[ijk], p[k], ex);
        if (!ssa.converged) atomicAdd(nonconv, 1);
        ql = ssa.ql; qi = ssa.qi;
    }
    out[ijk] = MODE == 1 ? ql : moist_buoyancy(ex, thl[ijk], qt[ijk], ql, qi, thvref[k]);
}

// get_buoyancy_surf (calc_buoyancy_bot, src/thermo_moist.cxx:637-655) and get_buoyancy_fluxbot (calc_buoyancy_fluxbot,
// :675-693): whole 2-D planes, ghost cells included; "assume no liquid water at the lowest model level".
// mode 0: b[kstart] and bbot from (thl, thlbot, qt, qtbot); mode 1: bfluxbot (in `bbot`) from the surface fluxes (in thlbot / qtbot).
template <typename TF>
__global__ void moist_surf_kernel(TF* __restrict__ b, TF* __restrict__ bbot, const TF* __restrict__ thl, const TF* __restrict__ thlbot,
        const TF* __restrict__ qt, const TF* __restrict__ qtbot, const TF* __restrict__ thvref, const TF* __restrict__ thvrefh,
        const GridDev<TF> g, const int mode)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.icells || j >= g.jcells) return;
    const long long ij = i + (long long)j * g.icells;
    const long long ijk = ij + g.kstart * g.ijcells;
    if (mode == 0)
    {
        bbot[ij] = moist_buoyancy_no_ql(thlbot[ij], qtbot[ij], thvrefh[g.kstart]);
        b[ijk] = moist_buoyancy_no_ql(thl[ijk], qt[ijk], thvref[g.kstart]);
    }
    else
        bbot[ij] = moist_buoyancy_flux_no_ql(thl[ijk], thlbot[ij], qt[ijk], qtbot[ij], thvrefh[g.kstart]);
}

#undef MHH_HD

} // namespace mhh
