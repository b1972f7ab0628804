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
// device: one reduction kernel for the two mean profiles and one single-CTA kernel for the hydrostatic integration (as a
// fixed-point sweep over all levels in parallel, see moist_base_state_kernel), both stream-ordered -- no host round trip, and
// the sub-step stays capturable in a CUDA graph.
// The buoyancy kernel is HBM-bound by design: thl and qt read once, wt read and written once = 4 array passes; threads march
// a few levels up their column with the next level's loads in flight; the exner function comes from the base state.
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

// Field3d_operators::calc_mean_profile for two fields at once: one CTA of 1024 threads per (level, field), double accumulation
// like the reference (its sum is sequential, this one a fixed-shape tree: deterministic, equal to rounding of the double sum).
// A warp walks whole rows (unit stride, no index division), four independent loads in flight per lane.
template <typename TF>
__global__ void __launch_bounds__(1024) moist_mean_profile_kernel(const TF* __restrict__ f0, const TF* __restrict__ f1,
        TF* __restrict__ m0, TF* __restrict__ m1, const GridDev<TF> g, const double n)
{
    const int k = blockIdx.x;
    const TF* __restrict__ f = (blockIdx.y == 0 ? f0 : f1) + (long long)k * g.ijcells + g.istart;
    TF* __restrict__ m = blockIdx.y == 0 ? m0 : m1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    double s0 = 0., s1 = 0., s2 = 0., s3 = 0.;
    for (int j = warp; j < g.jmax; j += nwarp)
    {
        const TF* __restrict__ row = f + (long long)(j + g.jstart) * g.icells;
        int i = lane;
        for (; i + 96 < g.imax; i += 128)
        {
            const TF a = row[i], b = row[i + 32], c = row[i + 64], d = row[i + 96];
            s0 += (double)a; s1 += (double)b; s2 += (double)c; s3 += (double)d;
        }
        for (; i < g.imax; i += 32) s0 += (double)row[i];
    }
    double s = (s0 + s1) + (s2 + s3);
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        double t = 0.;
        for (int w = 0; w < nwarp; ++w) t += red[w];
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

// ---- The same integration without the kmax-long chain of exp / pow / Newton calls on one lane ---------------------------
// (measured on the B200: 0.58 ms fp32 / 1.24 ms fp64 per launch at kmax = 256 for the serial form = 8 % of a 512 x 512 x 256 fp32
// step and 29 % of a 256^3 fp64 one; the reference's GPU build gave up on it, "extremely slow", src/thermo_moist.cu:917-923).
// The recurrences are  prefh[k+1] = prefh[k] F[k],  pref[k] = pref[k-1] Fh[k]  with  F[k] = exp(-g dz[k] / (Rd ex thv)(pref[k]))
// and  Fh[k] = exp(-g dzh[k] / (Rd exh thvh)(prefh[k]))  (z[kstart] for dzh at the surface).  Fixed-point form: from the current
// pressures ALL levels evaluate their factor in parallel (the expensive part: pow, exp, saturation adjustment), then one lane
// forms the two running products (kmax dependent multiplications, nothing else).  The dependence is strictly lower triangular
// -- level k needs only levels below it -- so after n sweeps the lowest n levels hold exactly the bits the serial code produces,
// and since the factor depends only weakly on the pressure (d ln F / d ln p ~ 1e-3 per level) all levels stop changing after a
// few sweeps: ~17 from scratch in fp64, 2-4 when the previous sub-step's pressures are the start.  The loop ends when a sweep
// changes no bit (or after kmax + 2 sweeps, the triangular bound), i.e. at THE fixed point, which is the serial result bit
// for bit -- tests/test_moist_hostcheck.py checks exactly that for the host build of these functions.
template <typename TF>
MHH_HD int moist_base_level(const MoistProfiles<TF> b, const TF* __restrict__ thlmean, const TF* __restrict__ qtmean, const int k,
        const int kstart, const int kend, const TF* __restrict__ z, const TF* __restrict__ dz, const TF* __restrict__ dzh,
        TF* __restrict__ F, TF* __restrict__ Fh)
{
    typedef MoistC<TF> C;
    int bad = 0;
    {
        const TF thli = TF(0.5) * (thlmean[k - 1] + thlmean[k]);
        const TF qti  = TF(0.5) * (qtmean[k - 1] + qtmean[k]);
        const TF ph = b.prefh[k];
        const TF exh = moist_exner(ph);
        const SatAdjust<TF> ssa = moist_sat_adjust(thli, qti, ph, exh);
        bad += ssa.converged ? 0 : 1;
        const TF thvh = moist_virtual_temperature(exh, thli, qti, ssa.ql, ssa.qi);
        b.exh[k] = exh; b.thvh[k] = thvh;
        b.rhoh[k] = ph / (C::Rd * exh * thvh);
        Fh[k] = m_exp(-C::grav * (k == kstart ? z[kstart] : dzh[k]) / (C::Rd * exh * thvh));
    }
    if (k < kend)
    {
        const TF p = b.pref[k];
        const TF ex = moist_exner(p);
        const SatAdjust<TF> ssa = moist_sat_adjust(thlmean[k], qtmean[k], p, ex);
        bad += ssa.converged ? 0 : 1;
        const TF thv = moist_virtual_temperature(ex, thlmean[k], qtmean[k], ssa.ql, ssa.qi);
        b.ex[k] = ex; b.thv[k] = thv;
        b.rho[k] = p / (C::Rd * ex * thv);
        F[k] = m_exp(-C::grav * dz[k] / (C::Rd * ex * thv));
    }
    return bad;
}

MHH_HD bool moist_same_bits(double a, double b) { return a == b || (a != a && b != b); }
MHH_HD bool moist_same_bits(float a, float b) { return a == b || (a != a && b != b); }

// the two running products; returns whether any pressure changed
template <typename TF>
MHH_HD int moist_base_cumprod(const MoistProfiles<TF> b, const TF pbot, const int kstart, const int kend,
        const TF* __restrict__ F, const TF* __restrict__ Fh)
{
    int changed = 0;
    TF ph = pbot, p = pbot;
    #pragma unroll 4
    for (int k = kstart; k <= kend; ++k)
    {
        // prefh[k] = prefh[k-1] * F[k-1]  (k > kstart);  pref[k] = pref[k-1] * Fh[k]  (pref[kstart] = prefh[kstart] * Fh[kstart])
        if (k > kstart) ph = ph * F[k - 1];
        p = (k == kstart ? ph : p) * Fh[k];
        changed |= (moist_same_bits(ph, b.prefh[k]) && moist_same_bits(p, b.pref[k])) ? 0 : 1;
        b.prefh[k] = ph; b.pref[k] = p;
    }
    return changed;
}

// One CTA.  cold != 0: start from p = pbot everywhere; else from the pressures already in b (the previous base state).
// info[0] += non-converged saturation adjustments of the final sweep, info[1] = number of sweeps.
template <typename TF>
__global__ void __launch_bounds__(256) moist_base_state_kernel(const MoistProfiles<TF> b, const TF* __restrict__ thlmean,
        const TF* __restrict__ qtmean, const TF pbot, const GridDev<TF> g, TF* __restrict__ F, TF* __restrict__ Fh,
        const int cold, int* __restrict__ info)
{
    __shared__ int s_changed, s_bad;
    const int kstart = g.kstart, kend = g.kend, tid = threadIdx.x, nth = blockDim.x;
    // no usable previous state (never computed, or profiles uploaded without pressures): start from scratch as well
    const bool scratch = cold != 0 || !(b.pref[kstart] > TF(0.)) || !(b.prefh[kend] > TF(0.));
    __syncthreads();
    if (scratch)
        for (int k = kstart + tid; k <= kend; k += nth) { b.prefh[k] = pbot; b.pref[k] = pbot; }
    __syncthreads();
    int sweeps = 0;
    const int maxsweeps = kend - kstart + 3;
    for (int it = 0; it < maxsweeps; ++it)
    {
        if (tid == 0) s_bad = 0;
        __syncthreads();
        int bad = 0;
        for (int k = kstart + tid; k <= kend; k += nth)
            bad += moist_base_level<TF>(b, thlmean, qtmean, k, kstart, kend, g.z, g.dz, g.dzh, F, Fh);
        if (bad) atomicAdd(&s_bad, bad);
        __syncthreads();
        if (tid == 0) s_changed = moist_base_cumprod<TF>(b, pbot, kstart, kend, F, Fh);
        __syncthreads();
        ++sweeps;
        if (!s_changed) break;
    }
    if (tid == 0)
    {
        b.pref[kstart - 1] = TF(2.) * b.prefh[kstart] - b.pref[kstart];
        if (s_bad) atomicAdd(&info[0], s_bad);
        info[1] = sweeps;
    }
}

// calc_buoyancy_tend_2nd (src/thermo_moist.cxx:77-120): wt += buoyancy of (thl, qt) interpolated to the half level, with the
// condensate of the saturation adjustment at that level's pressure.  One level per blockIdx.z (k = kstart+1 .. kend-1).
// exnh = exnrefh[k] of the base state: the reference evaluates exner(ph[k]) here, which is the very number its calc_base_state
// stored in exnrefh[k] (same function, same argument).
// A thread marches KCH levels up its column: thl and qt of the level below stay in registers (3 array reads per point instead
// of 4 out of L2) and the loads of the next level (thl, qt, wt) are issued before the arithmetic of the current one.  The
// one-point-per-thread version sat on two DRAM latencies in a row per 256-point CTA (loads -> arithmetic -> read-modify-write
// of wt): 0.61 ms per launch at 512 x 512 x 256 fp32 with no cloud at all (28 % of the HBM roofline; a pow per CTA in the very
// first version cost another 0.07 ms).
template <typename TF, int KCH>
__global__ void __launch_bounds__(256) moist_buoyancy_tend_kernel(TF* __restrict__ wt, const TF* __restrict__ thl, const TF* __restrict__ qt,
        const TF* __restrict__ ph, const TF* __restrict__ exh, const TF* __restrict__ thvrefh, const GridDev<TF> g, int* __restrict__ nonconv)
{
    const int i = g.istart + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.jstart + blockIdx.y * blockDim.y + threadIdx.y;
    const int k0 = g.kstart + 1 + blockIdx.z * KCH;
    const int k1 = min(k0 + KCH, g.kend);
    if (i >= g.iend || j >= g.jend || k0 >= k1) return;
    const long long kk = g.ijcells;
    long long ijk = i + (long long)j * g.icells + k0 * kk;
    TF thl_m = thl[ijk - kk], qt_m = qt[ijk - kk];
    TF thl_c = thl[ijk], qt_c = qt[ijk], wt_c = wt[ijk];
    int bad = 0;
    for (int k = k0; k < k1; ++k, ijk += kk)
    {
        TF thl_n = TF(0.), qt_n = TF(0.), wt_n = TF(0.);
        if (k + 1 < k1) { thl_n = thl[ijk + kk]; qt_n = qt[ijk + kk]; wt_n = wt[ijk + kk]; }
        const TF exnh = exh[k], p = ph[k];
        const TF thlh = interp2(thl_m, thl_c);
        const TF qth  = interp2(qt_m, qt_c);
        const SatAdjust<TF> ssa = moist_sat_adjust(thlh, qth, p, exnh);
        bad += ssa.converged ? 0 : 1;
        wt[ijk] = wt_c + moist_buoyancy(exnh, thlh, qth, ssa.ql, ssa.qi, thvrefh[k]);
        thl_m = thl_c; qt_m = qt_c;
        thl_c = thl_n; qt_c = qt_n; wt_c = wt_n;
    }
    if (bad) atomicAdd(nonconv, bad);
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
