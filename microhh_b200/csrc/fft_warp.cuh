// mhhb200 -- warp-per-sequence FFT kernels for Pres_2 (power-of-two lengths; the generic mixed-radix
// block kernels in poisson_kernels.cuh remain the fallback for 2^a 3^b 5^c sizes).
//
// One warp owns one sequence of L complex points in a warp-private, padded shared-memory row and runs
// an IN-PLACE decimation-in-frequency FFT with compile-time radices (8/4/2): no block barriers (only
// __syncwarp between stages), no index division, no radix dispatch, one buffer instead of a ping-pong
// pair.  The result is left in digit-reversed order; the consumers (real-FFT post-processing, the
// stores) read through digitrev(), which costs nothing extra.
//
// Reference behaviour restated (never copied): FFT<TF>::exec_forward/backward (src/fft.cxx:338-452,
// FFTW R2HC/HC2R semantics), Pres_2::input (src/pres_2.cxx:155-196), Pres_2::solve unpack (:332-361).
#pragma once
#include "common.cuh"
#include "poisson_kernels.cuh"

namespace mhh {

// padded index: breaks the power-of-two strides of the butterflies (16-byte elements, 8 per 128-byte bank row)
__host__ __device__ __forceinline__ constexpr int fpad(int i) { return i + (i >> 3) + (i >> 6); }
template <int L> struct FftRow { static constexpr int SIZE = fpad(L - 1) + 2; };

// radix plan of a length-L transform: up to 4 stages
template <int L> struct WPlan;
template <> struct WPlan<8>    { static constexpr int NS = 1, R0 = 8, R1 = 1, R2 = 1, R3 = 1; };
template <> struct WPlan<16>   { static constexpr int NS = 2, R0 = 4, R1 = 4, R2 = 1, R3 = 1; };
template <> struct WPlan<32>   { static constexpr int NS = 2, R0 = 8, R1 = 4, R2 = 1, R3 = 1; };
template <> struct WPlan<64>   { static constexpr int NS = 2, R0 = 8, R1 = 8, R2 = 1, R3 = 1; };
template <> struct WPlan<128>  { static constexpr int NS = 3, R0 = 8, R1 = 4, R2 = 4, R3 = 1; };
template <> struct WPlan<256>  { static constexpr int NS = 3, R0 = 8, R1 = 8, R2 = 4, R3 = 1; };
template <> struct WPlan<512>  { static constexpr int NS = 3, R0 = 8, R1 = 8, R2 = 8, R3 = 1; };
template <> struct WPlan<1024> { static constexpr int NS = 4, R0 = 8, R1 = 8, R2 = 4, R3 = 4; };
template <> struct WPlan<2048> { static constexpr int NS = 4, R0 = 8, R1 = 8, R2 = 8, R3 = 4; };

// position (in the in-place DIF result) of output frequency k
template <int L> __device__ __forceinline__ int digitrev(int k)
{
    typedef WPlan<L> P;
    // all radices and sub-lengths are compile-time powers of two: the divisions and remainders are shifts and masks
    int pos = (k % P::R0) * (L / P::R0); k /= P::R0;
    if (P::NS > 1) { pos += (k % P::R1) * (L / (P::R0 * P::R1)); k /= P::R1; }
    if (P::NS > 2) { pos += (k % P::R2) * (L / (P::R0 * P::R1 * P::R2)); k /= P::R2; }
    if (P::NS > 3) { pos += (k % P::R3) * (L / (P::R0 * P::R1 * P::R2 * P::R3)); }
    return pos;
}

// R-point DFT in registers (forward, e^{-2 pi i / R})
template <typename TF, int R> __device__ __forceinline__ void dft_regs(cplx<TF> (&a)[R])
{
    if (R == 2)
    {
        const cplx<TF> t = a[0];
        a[0] = cadd(t, a[1]); a[1] = csub(t, a[1]);
    }
    else if (R == 4)
    {
        const cplx<TF> t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]);
        const cplx<TF> t2 = cadd(a[1], a[3]), t3 = cmul_mi(csub(a[1], a[3]));
        a[0] = cadd(t0, t2); a[2] = csub(t0, t2);
        a[1] = cadd(t1, t3); a[3] = csub(t1, t3);
    }
    else if (R == 8)
    {
        const TF h = TF(0.70710678118654752440);
        cplx<TF> e[4], o[4];
        {
            const cplx<TF> t0 = cadd(a[0], a[4]), t1 = csub(a[0], a[4]);
            const cplx<TF> t2 = cadd(a[2], a[6]), t3 = cmul_mi(csub(a[2], a[6]));
            e[0] = cadd(t0, t2); e[2] = csub(t0, t2); e[1] = cadd(t1, t3); e[3] = csub(t1, t3);
        }
        {
            const cplx<TF> t0 = cadd(a[1], a[5]), t1 = csub(a[1], a[5]);
            const cplx<TF> t2 = cadd(a[3], a[7]), t3 = cmul_mi(csub(a[3], a[7]));
            o[0] = cadd(t0, t2); o[2] = csub(t0, t2); o[1] = cadd(t1, t3); o[3] = csub(t1, t3);
        }
        const cplx<TF> o1 = {h * (o[1].x + o[1].y), h * (o[1].y - o[1].x)};
        const cplx<TF> o2 = cmul_mi(o[2]);
        const cplx<TF> o3 = {h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y)};
        a[0] = cadd(e[0], o[0]); a[4] = csub(e[0], o[0]);
        a[1] = cadd(e[1], o1);   a[5] = csub(e[1], o1);
        a[2] = cadd(e[2], o2);   a[6] = csub(e[2], o2);
        a[3] = cadd(e[3], o3);   a[7] = csub(e[3], o3);
    }
}

// one in-place DIF stage: sub-transforms of length NCUR are split into R interleaved ones of length NCUR/R
template <typename TF, int L, int R, int NCUR>
__device__ __forceinline__ void wfft_stage(cplx<TF>* __restrict__ row, const cplx<TF>* __restrict__ tw, const int lane)
{
    constexpr int M = NCUR / R;
#pragma unroll
    for (int b = lane; b < L / R; b += 32)
    {
        const int blk = b / M, j = b % M;
        const int base = blk * NCUR + j;
        cplx<TF> a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = row[fpad(base + r * M)];
        dft_regs<TF, R>(a);
        row[fpad(base)] = a[0];
        if (M > 1)
        {
            // twiddle w_NCUR^(j r) = tw[j r L/NCUR], tw[t] = exp(-2 pi i t / L)
#pragma unroll
            for (int r = 1; r < R; ++r) row[fpad(base + r * M)] = cmul(a[r], tw[j * r * (L / NCUR)]);
        }
        else
        {
#pragma unroll
            for (int r = 1; r < R; ++r) row[fpad(base + r * M)] = a[r];
        }
    }
    __syncwarp();
}

// forward FFT of the warp's row (natural order in, digit-reversed order out)
template <typename TF, int L>
__device__ __forceinline__ void wfft(cplx<TF>* __restrict__ row, const cplx<TF>* __restrict__ tw, const int lane)
{
    typedef WPlan<L> P;
    constexpr int R0 = P::R0, R1 = P::R1, R2 = P::R2, R3 = P::R3;
    wfft_stage<TF, L, R0, L>(row, tw, lane);
    if (P::NS > 1) wfft_stage<TF, L, (R1 > 1 ? R1 : 2), (P::NS > 1 ? L / R0 : 2)>(row, tw, lane);
    if (P::NS > 2) wfft_stage<TF, L, (R2 > 1 ? R2 : 2), (P::NS > 2 ? L / (R0 * R1) : 2)>(row, tw, lane);
    if (P::NS > 3) wfft_stage<TF, L, (R3 > 1 ? R3 : 2), (P::NS > 3 ? L / (R0 * R1 * R2) : 2)>(row, tw, lane);
}

constexpr int WFFT_WARPS = 8;       // warps (= rows in flight) per CTA

template <typename TF, int L> constexpr size_t wfft_smem() { return (size_t)WFFT_WARPS * FftRow<L>::SIZE * sizeof(cplx<TF>); }

// ------------------------------------------------------------------------------------------
// x forward, fused with Pres_2::input: one warp per (j,k) row of itot = 2L reals.
// ------------------------------------------------------------------------------------------
template <typename TF, int L, bool RHS_FUSED>
__global__ void __launch_bounds__(32 * WFFT_WARPS) wfft_x_forward_kernel(TF* __restrict__ spec, const RhsSrc<TF> src, const GridDev<TF> g, const SpecLayout lay, const PeerPtrs<TF> pp,
        const cplx<TF>* __restrict__ tw_half, const cplx<TF>* __restrict__ tw_full, const long long nrows)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    cplx<TF>* row = reinterpret_cast<cplx<TF>*>(smem_raw) + warp * FftRow<L>::SIZE;
    constexpr int N = 2 * L, nm = L + 1;
    const long long jj = g.icells, kk = g.ijcells;
    const TF dti = src.dti, dxi = g.dxi, dyi = g.dyi;

    for (long long r = (long long)blockIdx.x * WFFT_WARPS + warp; r < nrows; r += (long long)gridDim.x * WFFT_WARPS)
    {
        if (RHS_FUSED)
        {
            const int kq = (int)(r / g.jmax);
            const int k = kq + g.kstart;
            const int j = (int)(r - (long long)kq * g.jmax) + g.jstart;
            const long long base = g.istart + j * jj + k * kk;
            const long long jn_off = (src.ywrap && j + 1 == g.jend) ? (1 - g.jmax) * jj : jj;
            const TF rho = g.rhoref[k], rhoh0 = g.rhorefh[k], rhoh1 = g.rhorefh[k + 1], dzi = g.dzi[k];
#pragma unroll 2
            for (int n = lane; n < L; n += 32)
            {
                const int i = 2 * n;
                const long long o = base + i;
                const long long o2 = (i + 2 == N) ? o + 2 - N : o + 2;          // periodic wrap instead of the ghost cell
                const TF u0 = src.ut[o] + src.u[o] * dti, u1 = src.ut[o + 1] + src.u[o + 1] * dti, u2 = src.ut[o2] + src.u[o2] * dti;
                const TF v0 = src.vt[o] + src.v[o] * dti, v1 = src.vt[o + 1] + src.v[o + 1] * dti;
                const TF vn0 = src.vt[o + jn_off] + src.v[o + jn_off] * dti, vn1 = src.vt[o + 1 + jn_off] + src.v[o + 1 + jn_off] * dti;
                const TF w0 = src.wt[o] + src.w[o] * dti, w1 = src.wt[o + 1] + src.w[o + 1] * dti;
                const TF wt0 = src.wt[o + kk] + src.w[o + kk] * dti, wt1 = src.wt[o + 1 + kk] + src.w[o + 1 + kk] * dti;
                cplx<TF> z;
                z.x = rho * (u1 - u0) * dxi + rho * (vn0 - v0) * dyi + (rhoh1 * wt0 - rhoh0 * w0) * dzi;
                z.y = rho * (u2 - u1) * dxi + rho * (vn1 - v1) * dyi + (rhoh1 * wt1 - rhoh0 * w1) * dzi;
                row[fpad(n)] = z;
            }
        }
        else
        {
            // rows of the workspace itself (pitch 2*nm reals), or of a separate compact (k, j, i) array (pitch itot) when src.u is set
            const cplx<TF>* in = src.u ? reinterpret_cast<const cplx<TF>*>(src.u + r * N) : reinterpret_cast<const cplx<TF>*>(spec + r * (2 * nm));
#pragma unroll
            for (int n = lane; n < L; n += 32) row[fpad(n)] = in[n];
        }
        __syncwarp();
        wfft<TF, L>(row, tw_half, lane);
        // real-FFT post-processing: X[m] = E[m] + W_N^m O[m], m = 0..L
        cplx<TF>* out = reinterpret_cast<cplx<TF>*>(spec) + (lay.P == 1 ? r * nm : 0);
#pragma unroll
        for (int m = lane; m < nm; m += 32)
        {
            const cplx<TF> zm = row[fpad(digitrev<L>(m == L ? 0 : m))];
            const cplx<TF> zc = cconj(row[fpad(digitrev<L>(m == 0 ? 0 : L - m))]);
            const cplx<TF> e = {TF(0.5) * (zm.x + zc.x), TF(0.5) * (zm.y + zc.y)};
            const cplx<TF> d = {TF(0.5) * (zm.x - zc.x), TF(0.5) * (zm.y - zc.y)};
            const cplx<TF> X = cadd(e, cmul(tw_full[m], cmul_mi(d)));
            if (pp.on) *peer_y_slot<TF>(pp, lay, r, m) = X;                  // straight into the owner's y-side buffer (NVLink store)
            else out[lay.P == 1 ? (long long)m : lay.xidx(r, m)] = X;
        }
        __syncwarp();
    }
    if (pp.on) __threadfence_system();          // peer stores are performed before the kernel counts as finished
}

// ------------------------------------------------------------------------------------------
// x backward fused with Pres_2::solve's unpack (ghost cells in x, y and the bottom level).
// ------------------------------------------------------------------------------------------
template <typename TF, int L>
__global__ void __launch_bounds__(32 * WFFT_WARPS) wfft_x_backward_kernel(const TF* __restrict__ spec, TF* __restrict__ p, const GridDev<TF> g, const SpecLayout lay,
        const cplx<TF>* __restrict__ tw_half, const cplx<TF>* __restrict__ tw_full, const long long nrows, const TF norm, const int fill_y_ghosts)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    cplx<TF>* row = reinterpret_cast<cplx<TF>*>(smem_raw) + warp * FftRow<L>::SIZE;
    constexpr int nm = L + 1;
    const long long jj = g.icells, kk = g.ijcells;

    for (long long r = (long long)blockIdx.x * WFFT_WARPS + warp; r < nrows; r += (long long)gridDim.x * WFFT_WARPS)
    {
        // Z'[m] = (X[m] + conj X[L-m]) + i e^{+2 pi i m/N} (X[m] - conj X[L-m]); conj(Z') goes in so that the forward
        // transform yields conj(inverse)
        const cplx<TF>* X = reinterpret_cast<const cplx<TF>*>(spec) + (lay.P == 1 ? r * nm : 0);
        const int kq0 = (int)(r / g.jmax), jq0 = (int)(r - (long long)kq0 * g.jmax);
#pragma unroll
        for (int m = lane; m < L; m += 32)
        {
            const cplx<TF> xm = X[lay.P == 1 ? (long long)m : lay.xidx_kj(kq0, jq0, m)];
            const cplx<TF> xc = cconj(X[lay.P == 1 ? (long long)(L - m) : lay.xidx_kj(kq0, jq0, L - m)]);
            const cplx<TF> e = cadd(xm, xc);
            const cplx<TF> d = csub(xm, xc);
            const cplx<TF> wd = cmul(cconj(tw_full[m]), d);
            row[fpad(m)] = {e.x - wd.y, -(e.y + wd.x)};
        }
        __syncwarp();
        wfft<TF, L>(row, tw_half, lane);
        const int kq = (int)(r / g.jmax);
        const int jq = (int)(r - (long long)kq * g.jmax);
        const long long rowbase = (jq + g.jstart) * jj + (kq + g.kstart) * kk;
        const bool ylo = fill_y_ghosts && jq < g.jgc, yhi = fill_y_ghosts && jq >= g.jmax - g.jgc;
        const int wtot = g.itot + 2 * g.igc;
        for (int ic = lane; ic < wtot; ic += 32)
        {
            int i = ic - g.igc;
            if (i < 0) i += g.itot; else if (i >= g.itot) i -= g.itot;
            const cplx<TF> zz = row[fpad(digitrev<L>(i >> 1))];
            const TF val = ((i & 1) ? -zz.y : zz.x) * norm;
            const long long o = ic + rowbase;
            p[o] = val;
            if (kq == 0) p[o - kk] = val;
            if (ylo) { p[o + g.jmax * jj] = val; if (kq == 0) p[o + g.jmax * jj - kk] = val; }
            if (yhi) { p[o - g.jmax * jj] = val; if (kq == 0) p[o - g.jmax * jj - kk] = val; }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// y transform (complex, length J): a CTA stages a panel of WFFT_WARPS consecutive x-modes (128 contiguous
// bytes per y in fp64) of one level, each warp transforms one mode, the panel goes back in place.
// ------------------------------------------------------------------------------------------
template <typename TF, int J>
__global__ void __launch_bounds__(32 * WFFT_WARPS) wfft_y_kernel(TF* __restrict__ spec, const SpecLayout lay, const PeerPtrs<TF> pp, const int nm, const int ktot,
        const cplx<TF>* __restrict__ tw, const int inverse)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int MC = WFFT_WARPS, NT = 32 * WFFT_WARPS, RS = FftRow<J>::SIZE;
    cplx<TF>* sm = reinterpret_cast<cplx<TF>*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npanel_m = (nm + MC - 1) / MC;
    const long long npanels = (long long)npanel_m * ktot;
    cplx<TF>* S = reinterpret_cast<cplx<TF>*>(spec);
    const int c = threadIdx.x % MC, j0 = threadIdx.x / MC;       // this thread's mode and first y within the panel

    for (long long pnl = blockIdx.x; pnl < npanels; pnl += gridDim.x)
    {
        const int k = (int)(pnl / npanel_m);
        const int m0 = (int)(pnl - (long long)k * npanel_m) * MC;
        const int mc = min(MC, nm - m0);
        cplx<TF>* base = S + (long long)k * J * nm + m0;      // P == 1 addressing; slabs go through lay.yidx
        const bool one = lay.P == 1;
        if (c < mc)
        {
#pragma unroll 4
            for (int j = j0; j < J; j += NT / MC)
            {
                cplx<TF> v = one ? base[(long long)j * nm + c] : S[lay.yidx(k, j, m0 + c)];
                if (inverse) v.y = -v.y;
                sm[c * RS + fpad(j)] = v;
            }
        }
        __syncthreads();
        if (warp < mc) wfft<TF, J>(sm + warp * RS, tw, lane);
        __syncthreads();
        if (c < mc)
        {
#pragma unroll 4
            for (int j = j0; j < J; j += NT / MC)
            {
                cplx<TF> v = sm[c * RS + fpad(digitrev<J>(j))];
                if (inverse) v.y = -v.y;
                if (one) base[(long long)j * nm + c] = v;
                else if (pp.on && inverse) *peer_x_slot<TF>(pp, lay, k, j, m0 + c) = v;       // straight into the row owner's x-side buffer
                else S[lay.yidx(k, j, m0 + c)] = v;
            }
        }
        __syncthreads();
    }
    if (pp.on && inverse) __threadfence_system();
}

} // namespace mhh
