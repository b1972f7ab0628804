// mhhb200 -- Pres_2 spectral solve, version 2: the y transforms FUSED with the Thomas sweeps.
//
//   x forward (rhs fused)            -> Y side, mode-major:  Y[src rank s][local mode ml][k][jl]
//   y forward + forward elimination  (one CTA per x-mode marches k upwards; in place, digit-reversed y order)
//   back substitution + y inverse    (one CTA per x-mode marches k downwards) -> X side: X[mode owner d][k][row panel][ml][8 rows]
//   x backward + unpack              -> ghosted p
//
// Seven array passes over the spectral workspace (1 W | 1 R + 1 W + table/2 | 1 R + table/2 + 1 W | 1 R) instead of the
// eleven of the three-kernel version (x forward | y forward | Thomas: 2 R + 2 W + table | y inverse | x backward); the
// tridiagonal coefficients never exist per point: the table holds winv[k] = 1 / (b[k] - a[k] c[k-1] winv[k-1]) for every
// (mode, level) -- it depends on the grid and the base state only -- so the sweeps are multiply-adds without a division.
//
// A y sequence (fixed x-mode, fixed level) is CONTIGUOUS in the mode-major layout: one warp loads it with 128-bit coalesced
// accesses into its private shared-memory row, transforms it there (in-place DIF, digit-reversed result) and applies the
// elimination in the digit-reversed order (the Thomas recurrence couples levels, not y-modes, so the order is immaterial);
// consecutive levels are pipelined over the warps of the CTA, the only coupling being the previous level's solution handed
// over in a two-slot shared-memory ring behind a "levels done" counter.  The downward pass mirrors it and ends with the matching
// decimation-in-time inverse (digit-reversed in, natural out), so no permutation pass exists anywhere.
//
// Both transposes of the y-slab decomposition are the store phases of the x-forward and y-inverse kernels (8-row chunks of
// 128 bytes into the owning rank's buffer over NVLink peer mappings, or into a staging buffer whose blocks ARE the
// ncclSend messages).  Reference semantics: Pres_2::solve + tdma (src/pres_2.cxx:202-361), FFT::exec_forward/backward
// (src/fft.cxx:338-452 serial, :455-587 MPI), Transpose::exec_xy/yx (src/transpose.cxx:117-271).
#pragma once
#include "fft_warp.cuh"
#include "tile3_kernels.cuh"      // mbarrier primitives, cp_async

namespace mhh {

constexpr int P2_ROWS = 8;        // rows per x panel (= warps per CTA of the x kernels; 8 x 16 B = one 128-byte store chunk)

// Layout of the two spectral buffers (complex elements).
struct Spec2
{
    int P, rank;
    int nm, base, rem;        // x-modes: nm = itot/2 + 1, dealt out base (+1 for the first rem ranks)
    int jmax, jtot, ktot;
    int mcl, m_off;           // this rank's modes
    int npan;                 // row panels per level: ceil(jmax / 8)

    __host__ __device__ int count(const int d) const { return base + (d < rem ? 1 : 0); }
    __host__ __device__ int offset(const int d) const { return d * base + (d < rem ? d : rem); }
    __host__ __device__ int owner(const int m) const
    {
        const int cut = rem * (base + 1);
        return m < cut ? m / (base + 1) : rem + (m - cut) / base;
    }
    // Y side of a rank owning `cnt` modes: element (source rank s, local mode ml, level k, local row jl)
    __host__ __device__ long long yidx(const int cnt, const int s, const int ml, const int k, const int jl) const
    { return (((long long)s * cnt + ml) * ktot + k) * jmax + jl; }
    // X side of any rank: element (mode owner d, level k, row panel jp, local mode ml of d, row r of the panel)
    __host__ __device__ long long xidx(const int d, const int k, const int jp, const int ml, const int r) const
    { return 8 * ((long long)offset(d) * ktot * npan + ((long long)k * npan + jp) * count(d) + ml) + r; }
    __host__ __device__ long long yside_elems() const { return (long long)mcl * ktot * jtot; }
    __host__ __device__ long long xside_elems() const { return 8LL * nm * ktot * npan; }
};

inline Spec2 make_spec2(int itot, int jtot, int ktot, int P, int rank)
{
    Spec2 s{};
    s.P = P; s.rank = rank; s.nm = itot / 2 + 1;
    s.base = s.nm / P; s.rem = s.nm % P;
    s.jmax = jtot / P; s.jtot = jtot; s.ktot = ktot;
    s.mcl = s.count(rank); s.m_off = s.offset(rank);
    s.npan = (s.jmax + P2_ROWS - 1) / P2_ROWS;
    return s;
}

// where the store phases of the two transposing kernels write: the owner's buffer (peer mapping or own), or the local staging
// buffer of the NCCL transport (dst[d] = local base of the block for rank d, laid out like the destination block)
template <typename TF>
struct XferPtrs
{
    cplx<TF>* dst[MAX_SLAB_RANKS];
    int staged;       // 1: dst[d] is a private staging block (indexing relative to the block), 0: dst[d] is rank d's whole buffer
};

// ---- inverse (decimation in time) stages: exact inverses of wfft_stage, applied in reverse order ----------------------
template <typename TF, int R> __device__ __forceinline__ void idft_regs(cplx<TF> (&a)[R])
{
    // IDFT(a) = swap(DFT(swap(a))), swap = exchange of real and imaginary parts
#pragma unroll
    for (int r = 0; r < R; ++r) { const TF t = a[r].x; a[r].x = a[r].y; a[r].y = t; }
    dft_regs<TF, R>(a);
#pragma unroll
    for (int r = 0; r < R; ++r) { const TF t = a[r].x; a[r].x = a[r].y; a[r].y = t; }
}

template <typename TF, int L, int R, int NCUR>
__device__ __forceinline__ void wfft_inv_stage(cplx<TF>* __restrict__ row, const cplx<TF>* __restrict__ tw, const int lane)
{
    constexpr int M = NCUR / R;
#pragma unroll
    for (int b = lane; b < L / R; b += 32)
    {
        const int blk = b / M, j = b % M;
        const int base = blk * NCUR + j;
        cplx<TF> a[R];
        a[0] = row[fpad(base)];
        if (M > 1)
        {
#pragma unroll
            for (int r = 1; r < R; ++r) a[r] = cmul(row[fpad(base + r * M)], cconj(tw[j * r * (L / NCUR)]));
        }
        else
        {
#pragma unroll
            for (int r = 1; r < R; ++r) a[r] = row[fpad(base + r * M)];
        }
        idft_regs<TF, R>(a);
#pragma unroll
        for (int r = 0; r < R; ++r) row[fpad(base + r * M)] = a[r];
    }
    __syncwarp();
}

// unnormalised inverse FFT of the warp's row: digit-reversed order in (as wfft leaves it), natural order out, times L
template <typename TF, int L>
__device__ __forceinline__ void wfft_inv(cplx<TF>* __restrict__ row, const cplx<TF>* __restrict__ tw, const int lane)
{
    typedef WPlan<L> P;
    constexpr int R0 = P::R0, R1 = P::R1, R2 = P::R2, R3 = P::R3;
    if (P::NS > 3) wfft_inv_stage<TF, L, (R3 > 1 ? R3 : 2), (P::NS > 3 ? L / (R0 * R1 * R2) : 2)>(row, tw, lane);
    if (P::NS > 2) wfft_inv_stage<TF, L, (R2 > 1 ? R2 : 2), (P::NS > 2 ? L / (R0 * R1) : 2)>(row, tw, lane);
    if (P::NS > 1) wfft_inv_stage<TF, L, (R1 > 1 ? R1 : 2), (P::NS > 1 ? L / R0 : 2)>(row, tw, lane);
    wfft_inv_stage<TF, L, R0, L>(row, tw, lane);
}

// frequency held at position `pos` of the digit-reversed result (inverse of digitrev<L>)
template <int L> __host__ __device__ __forceinline__ int digitrev_inv(int pos)
{
    typedef WPlan<L> P;
    int k = 0, w = 1;
    { const int n = L / P::R0; k += (pos / n) * w; pos %= n; w *= P::R0; }
    if (P::NS > 1) { const int n = L / (P::R0 * P::R1); k += (pos / n) * w; pos %= n; w *= P::R1; }
    if (P::NS > 2) { const int n = L / (P::R0 * P::R1 * P::R2); k += (pos / n) * w; pos %= n; w *= P::R2; }
    if (P::NS > 3) { k += pos * w; }
    return k;
}

// ------------------------------------------------------------------------------------------
// Two-sided (twisted) factorisation of the tridiagonal systems of Pres_2::solve (src/pres_2.cxx:202-324): the lower half of
// the levels, 0 .. ksplit-1, is eliminated upwards exactly like the reference's tdma (pivots w_k), the upper half,
// ktot-1 .. ksplit, downwards (pivots v_k); the two meet in a 2 x 2 system per column.  Mathematically the same solution (the
// rounding differs from a one-sided sweep at the 1e-16 level); the point is that the two halves are independent sequential
// chains, so a mode is worked on by TWO CTAs at a time -- twice the parallelism where modes are scarce (y slabs: nm / P modes
// per GPU) and half the length of the critical path.
//   lower: w_0 = b_0,  w_k = b_k - a_k c_{k-1} / w_{k-1};   p'_k = (d_k - a_k p'_{k-1}) / w_k;   x_k = p'_k - (c_k / w_k) x_{k+1}
//   upper: v_K = b_K,  v_k = b_k - c_k a_{k+1} / v_{k+1};   q'_k = (d_k - c_k q'_{k+1}) / v_k;   x_k = q'_k - (a_k / v_k) x_{k-1}
//   interface (m = ksplit):  x_m = (q'_m - beta p'_{m-1}) / (1 - alpha beta),  x_{m-1} = p'_{m-1} - alpha x_m,
//                            alpha = c_{m-1} / w_{m-1},  beta = a_m / v_m
// Table T[ml][k][pos] = 1 / w_k (k < ksplit) or 1 / v_k (k >= ksplit) of column (mode m_off + ml, y-mode at digit-reversed
// position pos); Dinv[ml][pos] = 1 / (1 - alpha beta) follows the table.
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int p2_ksplit(const int ktot) { return ktot / 2; }

template <typename TF, int J>
__global__ void tdma2_setup_kernel(TF* __restrict__ T, TF* __restrict__ Dinv, const TdmaCoef<TF> cf, const int mcl, const int kmax, const int m_off)
{
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= (long long)mcl * J) return;
    const int ml = (int)(col / J), pos = (int)(col - (long long)ml * J);
    const int l = digitrev_inv<J>(pos), m = ml + m_off;
    const TF lam = cf.bmati[m] + cf.bmatj[l];
    const bool mode00 = (m == 0) && (l == 0);
    const int ks = p2_ksplit(kmax);
    TF* t = T + (long long)ml * kmax * J + pos;
    TF w = tdma_b(cf, 0, kmax, lam, mode00);
    t[0] = TF(1) / w;
    for (int k = 1; k < ks; ++k)
    {
        const TF f = cf.c[k - 1] / w;
        w = tdma_b(cf, k, kmax, lam, mode00) - cf.a[k] * f;
        t[(long long)k * J] = TF(1) / w;
    }
    TF v = tdma_b(cf, kmax - 1, kmax, lam, mode00);
    t[(long long)(kmax - 1) * J] = TF(1) / v;
    for (int k = kmax - 2; k >= ks; --k)
    {
        const TF f = cf.a[k + 1] / v;
        v = tdma_b(cf, k, kmax, lam, mode00) - cf.c[k] * f;
        t[(long long)k * J] = TF(1) / v;
    }
    const TF alpha = cf.c[ks - 1] / w, beta = cf.a[ks] / v;
    Dinv[col] = TF(1) / (TF(1) - alpha * beta);
}

// ------------------------------------------------------------------------------------------
// x forward, fused with Pres_2::input: a CTA owns a panel of 8 consecutive rows of one level (one warp per row), then
// stores mode by mode: the 8 rows of a mode are one 128-byte chunk of the mode owner's Y side.
// RHS = 0: the rows come from a compact (k, j, i) real array (tests, Pres_4); 1: right-hand side fused, scalar loads;
// 2: fused with one vector load per x-pair (needs an even istart / icells and vector-aligned arrays: 14 loads per two
// points instead of 22, twice the bytes in flight per load -- the kernel waits on global-load latency).
// ------------------------------------------------------------------------------------------
// (L = 512, itot = 1024: the unrolled transform takes 110 registers in fp64, two CTAs per SM; capping it to 80 for three was
// measured at 1024^3: 52.4 -> 77.4 ms/step, the spills cost more than the occupancy gains)
template <typename TF, int L, int RHS>
__global__ void __launch_bounds__(32 * P2_ROWS) p2_x_forward_kernel(const TF* __restrict__ compact, const RhsSrc<TF> src, const GridDev<TF> g,
        const Spec2 lay, const XferPtrs<TF> xf, const cplx<TF>* __restrict__ tw_half, const cplx<TF>* __restrict__ tw_full)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int RS = FftRow<L>::SIZE;
    cplx<TF>* rows = reinterpret_cast<cplx<TF>*>(smem_raw);
    cplx<TF>* row = rows + warp * RS;
    constexpr int N = 2 * L, nm = L + 1;
    const long long jj = g.icells, kk = g.ijcells;
    const TF dti = src.dti, dxi = g.dxi, dyi = g.dyi;
    const int npanels = lay.npan * g.ktot;

    for (int pnl = blockIdx.x; pnl < npanels; pnl += gridDim.x)
    {
        const int kq = pnl / lay.npan, jp = pnl - kq * lay.npan;
        const int jl = jp * P2_ROWS + warp;
        if (jl < g.jmax)
        {
            if (RHS == 2)
            {
                typedef typename V2T<TF>::type V2;
                const int k = kq + g.kstart;
                const int j = jl + g.jstart;
                const long long base = g.istart + j * jj + k * kk;
                const long long jn_off = (src.ywrap && j + 1 == g.jend) ? (1 - g.jmax) * jj : jj;
                const TF rho = g.rhoref[k], rhoh0 = g.rhorefh[k], rhoh1 = g.rhorefh[k + 1], dzi = g.dzi[k];
                auto LD = [](const TF* p) -> V2 { return *reinterpret_cast<const V2*>(p); };
#pragma unroll 2
                for (int n = lane; n < L; n += 32)
                {
                    const int i = 2 * n;
                    const long long o = base + i;
                    const long long o2 = (i + 2 == N) ? o + 2 - N : o + 2;          // periodic wrap instead of the ghost cell
                    const V2 ut = LD(src.ut + o), uu = LD(src.u + o);
                    const V2 vt = LD(src.vt + o), vv = LD(src.v + o), vtn = LD(src.vt + o + jn_off), vvn = LD(src.v + o + jn_off);
                    const V2 wt = LD(src.wt + o), ww = LD(src.w + o), wtt = LD(src.wt + o + kk), wwt = LD(src.w + o + kk);
                    const TF u0 = ut.x + uu.x * dti, u1 = ut.y + uu.y * dti, u2 = src.ut[o2] + src.u[o2] * dti;
                    const TF v0 = vt.x + vv.x * dti, v1 = vt.y + vv.y * dti;
                    const TF vn0 = vtn.x + vvn.x * dti, vn1 = vtn.y + vvn.y * dti;
                    const TF w0 = wt.x + ww.x * dti, w1 = wt.y + ww.y * dti;
                    const TF wt0 = wtt.x + wwt.x * dti, wt1 = wtt.y + wwt.y * dti;
                    cplx<TF> z;
                    z.x = rho * (u1 - u0) * dxi + rho * (vn0 - v0) * dyi + (rhoh1 * wt0 - rhoh0 * w0) * dzi;
                    z.y = rho * (u2 - u1) * dxi + rho * (vn1 - v1) * dyi + (rhoh1 * wt1 - rhoh0 * w1) * dzi;
                    row[fpad(n)] = z;
                }
            }
            else if (RHS == 1)
            {
                const int k = kq + g.kstart;
                const int j = jl + g.jstart;
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
                const cplx<TF>* in = reinterpret_cast<const cplx<TF>*>(compact + ((long long)kq * g.jmax + jl) * N);
#pragma unroll
                for (int n = lane; n < L; n += 32) row[fpad(n)] = in[n];
            }
            __syncwarp();
            wfft<TF, L>(row, tw_half, lane);
        }
        __syncthreads();
        // real-FFT post-processing X[m] = E[m] + W_N^m O[m], m = 0..L, and the transposing store: thread -> (mode, row)
        const int r = threadIdx.x & (P2_ROWS - 1);
        const bool rvalid = jp * P2_ROWS + r < g.jmax;
        const cplx<TF>* rr = rows + r * RS;
        for (int m = threadIdx.x / P2_ROWS; m < nm; m += 32)
        {
            if (!rvalid) continue;
            const cplx<TF> zm = rr[fpad(digitrev<L>(m == L ? 0 : m))];
            const cplx<TF> zc = cconj(rr[fpad(digitrev<L>(m == 0 ? 0 : L - m))]);
            const cplx<TF> e = {TF(0.5) * (zm.x + zc.x), TF(0.5) * (zm.y + zc.y)};
            const cplx<TF> d = {TF(0.5) * (zm.x - zc.x), TF(0.5) * (zm.y - zc.y)};
            const cplx<TF> X = cadd(e, cmul(tw_full[m], cmul_mi(d)));
            const int dd = lay.P == 1 ? 0 : lay.owner(m);
            const int cnt = lay.count(dd), ml = m - lay.offset(dd);
            // staged: the block for rank dd is laid out like block `rank` of dd's Y side, i.e. [ml][k][jl]
            const long long idx = lay.yidx(cnt, xf.staged ? 0 : lay.rank, ml, kq, jp * P2_ROWS + r);
            xf.dst[dd][idx] = X;
        }
        __syncthreads();
    }
    if (lay.P > 1 && !xf.staged) __threadfence_system();          // peer stores are performed before the kernel counts as finished
}

// ------------------------------------------------------------------------------------------
// y forward transform + forward elimination, one CTA per local x-mode (blockIdx.x), warps pipeline the levels upwards.
//   Y[.][ml][k][.] (natural y order)  ->  p'_k in digit-reversed y order, in place.   solve = 0: transform only.
// Shared memory: [level counter: 128 B][state ring: 2 x J][rows: NW x FftRow<J>]
// Hand-over protocol: level k writes ring slot k & 1 after it has read slot (k-1) & 1, then publishes done = k + 1;
// level k+1 waits for done >= k + 1.  Level k+2 reuses slot k & 1 only after level k+1 -- the slot's only reader -- has
// published, so two slots suffice and no "slot free" signal is needed.
// ------------------------------------------------------------------------------------------
template <typename TF, int J> constexpr int p2_y_warps() { return (sizeof(TF) * J > 8 * 1024) ? 4 : 8; }     // fp64 J = 2048: 4 rows fit
template <typename TF, int J> constexpr size_t p2_y_smem()
{ return 128 + ((size_t)2 * J + (size_t)p2_y_warps<TF, J>() * FftRow<J>::SIZE) * sizeof(cplx<TF>); }

// linear y index (natural or digit-reversed position) -> offset inside a (mode, level) sequence of the Y side
__device__ __forceinline__ long long p2_yoff(const Spec2& lay, const int jlog2, const int j)
{
    if (lay.P == 1) return j;
    const int s = j >> jlog2, jl = j & (lay.jmax - 1);
    return (long long)s * lay.mcl * lay.ktot * lay.jmax + jl;
}

// Level hand-over between the warps of a CTA: a monotonically increasing "levels done" counter in shared memory.
// (mbarrier parities cannot be used here: a warp may start waiting several levels ahead of the producer, and a parity wait
// that starts more than one phase early returns at the wrong completion.)
__device__ __forceinline__ void level_wait(const volatile int* done, const int need)
{
    while (*done < need) __nanosleep(40);
    __threadfence_block();          // acquire: the producer's ring slot is visible
}
__device__ __forceinline__ void level_publish(volatile int* done, const int value, const int lane)
{
    __syncwarp();
    if (lane == 0) { __threadfence_block(); *done = value; }
}

__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <typename TF, int J>
__global__ void __launch_bounds__(32 * p2_y_warps<TF, J>(), (sizeof(TF) * J <= 8 * 512) ? 2 : 1) p2_y_forward_kernel(cplx<TF>* __restrict__ Y, const TF* __restrict__ T, const Spec2 lay,
        const TF* __restrict__ ak, const TF* __restrict__ ck, const TF* __restrict__ dz2, const cplx<TF>* __restrict__ tw, const int jlog2, const int solve)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = p2_y_warps<TF, J>(), RS = FftRow<J>::SIZE, NI = (J + 31) / 32;
    constexpr bool TREG = NI <= 32;          // table entries of the level prefetched into registers (issued before the transform)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    cplx<TF>* state = reinterpret_cast<cplx<TF>*>(smem_raw + 128);
    cplx<TF>* row = state + 2 * J + warp * RS;
    volatile int* done = reinterpret_cast<volatile int*>(smem_raw);          // levels completed so far (in order)
    if (threadIdx.x == 0) *done = 0;
    __syncthreads();
    const int K = lay.ktot;
    // blockIdx.x = 2 * mode + half: the lower half eliminates levels 0 .. ksplit-1 upwards (coupling a_k to the level below),
    // the upper half levels ktot-1 .. ksplit downwards (coupling c_k to the level above)
    const int ml = blockIdx.x >> 1, half = blockIdx.x & 1;
    const int ksp = p2_ksplit(K);
    const int nlev = half ? K - ksp : ksp;
    const TF* __restrict__ coef = half ? ck : ak;
    cplx<TF>* Ym = Y + (long long)ml * K * lay.jmax;
    const TF* Tm = T + (long long)ml * K * J;
    for (int n = warp; n < nlev; n += NW)
    {
        const int k = half ? K - 1 - n : n;
        cplx<TF>* seq = Ym + (long long)k * lay.jmax;
        TF tk[TREG ? NI : 1];
        if (solve && TREG)
        {
#pragma unroll
            for (int i = 0; i < NI; ++i) { const int pos = lane + 32 * i; tk[i] = pos < J ? Tm[(long long)k * J + pos] : TF(0); }
        }
#pragma unroll 4
        for (int j = lane; j < J; j += 32) row[fpad(j)] = seq[p2_yoff(lay, jlog2, j)];
        __syncwarp();
        wfft<TF, J>(row, tw, lane);
        if (solve)
        {
            const TF a = coef[k], d2 = dz2[k];
            if (n > 0) level_wait(done, n);                     // the previous level of this half has published its slot
            const cplx<TF>* prev = state + ((n - 1) & 1) * J;
            cplx<TF>* cur = state + (n & 1) * J;
#pragma unroll
            for (int i = 0; i < NI; ++i)
            {
                const int pos = lane + 32 * i;
                if (pos < J)
                {
                    const TF t = TREG ? tk[TREG ? i : 0] : Tm[(long long)k * J + pos];
                    const cplx<TF> z = row[fpad(pos)];
                    cplx<TF> pv = {TF(0), TF(0)};
                    if (n > 0) pv = prev[pos];
                    const cplx<TF> o = {(d2 * z.x - a * pv.x) * t, (d2 * z.y - a * pv.y) * t};
                    cur[pos] = o;
                    seq[p2_yoff(lay, jlog2, pos)] = o;
                }
            }
            level_publish(done, n + 1, lane);
        }
        else
        {
#pragma unroll 4
            for (int pos = lane; pos < J; pos += 32) seq[p2_yoff(lay, jlog2, pos)] = row[fpad(pos)];
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// back substitution + inverse y transform, one CTA per local x-mode, warps pipeline the levels downwards; the result
// (natural y order, unnormalised) goes to the X side of the rank that owns the rows, in 8-row chunks.
// ------------------------------------------------------------------------------------------
template <typename TF, int J>
__global__ void __launch_bounds__(32 * p2_y_warps<TF, J>(), (sizeof(TF) * J <= 8 * 512) ? 2 : 1) p2_y_backward_kernel(const cplx<TF>* __restrict__ Y, const TF* __restrict__ T, const TF* __restrict__ Dinv,
        const Spec2 lay, const XferPtrs<TF> xf, const TF* __restrict__ ak, const TF* __restrict__ ck, const cplx<TF>* __restrict__ tw, const int jlog2, const int solve)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = p2_y_warps<TF, J>(), RS = FftRow<J>::SIZE, NI = (J + 31) / 32;
    constexpr bool TREG = NI <= 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    cplx<TF>* state = reinterpret_cast<cplx<TF>*>(smem_raw + 128);
    cplx<TF>* row = state + 2 * J + warp * RS;
    volatile int* done = reinterpret_cast<volatile int*>(smem_raw);
    if (threadIdx.x == 0) *done = 0;
    __syncthreads();
    const int K = lay.ktot;
    const int cnt = lay.mcl;
    // blockIdx.x = 2 * mode + half: the lower half substitutes from level ksplit-1 down to 0, the upper half from ksplit up to
    // ktot-1; both start from the 2 x 2 interface system (each solves it for itself: two loads and five flops per column)
    const int ml = blockIdx.x >> 1, half = blockIdx.x & 1;
    const int ksp = p2_ksplit(K);
    const int nlev = half ? K - ksp : ksp;
    const TF* __restrict__ coef = half ? ak : ck;
    const cplx<TF>* Ym = Y + (long long)ml * K * lay.jmax;
    const TF* Tm = T + (long long)ml * K * J;
    for (int n = warp; n < nlev; n += NW)
    {
        const int k = half ? ksp + n : ksp - 1 - n;
        const cplx<TF>* seq = Ym + (long long)k * lay.jmax;
        // p'_k straight into the warp's row (asynchronous copies: issued before the hand-over wait, no registers)
#pragma unroll 4
        for (int pos = lane; pos < J; pos += 32) cp_async<(int)sizeof(cplx<TF>)>(&row[fpad(pos)], &seq[p2_yoff(lay, jlog2, pos)]);
        if (solve)
        {
            // the coupling factor of this level: c_k / w_k (lower half) or a_k / v_k (upper half) = coef_k * T[k]
            TF fk[TREG ? NI : 1];
            const TF c = coef[k];
            if (TREG)
            {
#pragma unroll
                for (int i = 0; i < NI; ++i) { const int pos = lane + 32 * i; fk[i] = pos < J ? c * Tm[(long long)k * J + pos] : TF(0); }
            }
            cp_async_wait_all();
            __syncwarp();
            if (n > 0)
            {
                level_wait(done, n);
                const cplx<TF>* prev = state + ((n - 1) & 1) * J;
                cplx<TF>* cur = state + (n & 1) * J;
#pragma unroll
                for (int i = 0; i < NI; ++i)
                {
                    const int pos = lane + 32 * i;
                    if (pos < J)
                    {
                        const TF f = TREG ? fk[TREG ? i : 0] : c * Tm[(long long)k * J + pos];
                        cplx<TF> o = row[fpad(pos)];
                        const cplx<TF> nx = prev[pos];
                        o.x -= f * nx.x; o.y -= f * nx.y;
                        cur[pos] = o;
                        row[fpad(pos)] = o;
                    }
                }
            }
            else
            {
                // first level of the half: the interface.  x_m = (q'_m - beta p'_{m-1}) Dinv;  lower half: x_{m-1} = p'_{m-1} - alpha x_m
                const cplx<TF>* oth = Ym + (long long)(half ? ksp - 1 : ksp) * lay.jmax;      // the other half's last eliminated level
                const TF* To = Tm + (long long)(half ? ksp - 1 : ksp) * J;
                const TF co = half ? ck[ksp - 1] : ak[ksp];
                cplx<TF>* cur = state;
#pragma unroll 4
                for (int pos = lane; pos < J; pos += 32)
                {
                    const cplx<TF> own = row[fpad(pos)];
                    const cplx<TF> ot = oth[p2_yoff(lay, jlog2, pos)];
                    const TF fo = c * Tm[(long long)k * J + pos];          // own coupling factor: alpha (lower) or beta (upper)
                    const TF fx = co * To[pos];                            // the other one
                    const TF di = Dinv[(long long)ml * J + pos];
                    const TF beta = half ? fo : fx, alpha = half ? fx : fo;
                    const cplx<TF> q = half ? own : ot, pm = half ? ot : own;   // q'_m, p'_{m-1}
                    const cplx<TF> xm = {(q.x - beta * pm.x) * di, (q.y - beta * pm.y) * di};
                    cplx<TF> o = xm;
                    if (!half) { o.x = pm.x - alpha * xm.x; o.y = pm.y - alpha * xm.y; }
                    cur[pos] = o;
                    row[fpad(pos)] = o;
                }
            }
            level_publish(done, n + 1, lane);
        }
        else
        {
            cp_async_wait_all();
            __syncwarp();
        }
        wfft_inv<TF, J>(row, tw, lane);
        // natural y order -> 8-row chunks of the row owner's X side
#pragma unroll 4
        for (int j = lane; j < J; j += 32)
        {
            const int s = lay.P == 1 ? 0 : (j >> jlog2);
            const int jl = lay.P == 1 ? j : (j & (lay.jmax - 1));
            const int jp = jl / P2_ROWS, r = jl % P2_ROWS;
            long long idx;
            if (xf.staged) idx = 8 * (((long long)k * lay.npan + jp) * cnt + ml) + r;          // private block for rank s: [k][jp][ml][8]
            else idx = lay.xidx(lay.rank, k, jp, ml, r);
            xf.dst[s][idx] = row[fpad(j)];
        }
        __syncwarp();
    }
    if (lay.P > 1 && !xf.staged) __threadfence_system();
}

// ------------------------------------------------------------------------------------------
// x backward fused with Pres_2::solve's unpack: a CTA owns a row panel (8 rows of one level); the panel's modes are
// contiguous per mode owner.  Writes the ghosted p including the periodic x ghosts, the bottom ghost level and (single
// GPU) the periodic y ghosts.
// ------------------------------------------------------------------------------------------
template <typename TF, int L>
__global__ void __launch_bounds__(32 * P2_ROWS, 3) p2_x_backward_kernel(const cplx<TF>* __restrict__ X, TF* __restrict__ p, const GridDev<TF> g, const Spec2 lay,
        const cplx<TF>* __restrict__ tw_half, const cplx<TF>* __restrict__ tw_full, const TF norm, const int fill_y_ghosts)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int RS = FftRow<L>::SIZE;
    cplx<TF>* rows = reinterpret_cast<cplx<TF>*>(smem_raw);
    cplx<TF>* row = rows + warp * RS;
    const long long jj = g.icells, kk = g.ijcells;
    const int npanels = lay.npan * g.ktot;

    for (int pnl = blockIdx.x; pnl < npanels; pnl += gridDim.x)
    {
        const int kq = pnl / lay.npan, jp = pnl - kq * lay.npan;
        // Z'[m] = (X[m] + conj X[L-m]) + i e^{+2 pi i m/N} (X[m] - conj X[L-m]); conj(Z') goes in so that the forward
        // transform yields conj(inverse).  Thread -> (mode pair {m, L-m}, row): every element is read once.
        {
            const int r = threadIdx.x & (P2_ROWS - 1);
            cplx<TF>* rr = rows + r * RS;
            auto ld = [&](const int m) -> cplx<TF> {
                const int dd = lay.P == 1 ? 0 : lay.owner(m);
                return X[lay.xidx(dd, kq, jp, m - lay.offset(dd), r)];
            };
            for (int m = threadIdx.x / P2_ROWS; m <= L / 2; m += 32)
            {
                const int mc = L - m;
                const cplx<TF> xa = ld(m), xb = ld(mc);
                {
                    const cplx<TF> xc = cconj(xb);
                    const cplx<TF> e = cadd(xa, xc), d = csub(xa, xc);
                    const cplx<TF> wd = cmul(cconj(tw_full[m]), d);
                    rr[fpad(m)] = {e.x - wd.y, -(e.y + wd.x)};
                }
                if (m != 0 && mc != m)
                {
                    const cplx<TF> xc = cconj(xa);
                    const cplx<TF> e = cadd(xb, xc), d = csub(xb, xc);
                    const cplx<TF> wd = cmul(cconj(tw_full[mc]), d);
                    rr[fpad(mc)] = {e.x - wd.y, -(e.y + wd.x)};
                }
            }
        }
        __syncthreads();
        const int jq = jp * P2_ROWS + warp;
        if (jq < g.jmax)
        {
            wfft<TF, L>(row, tw_half, lane);
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
        }
        __syncthreads();
    }
}

} // namespace mhh
