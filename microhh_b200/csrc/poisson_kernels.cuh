// mhhb200 -- Poisson solver kernels for Pres_2: shared-memory Stockham FFTs (mixed radix
// 2/3/4/5/8) in x (real <-> half-spectrum) and y (complex), and the per-mode tridiagonal
// solve in z.  cuFFT is NOT used here; it is only a comparator in the tests.
//
// Reference behaviour restated (never copied): FFT<TF>::exec_forward/backward (src/fft.cxx:338-452,
// FFTW r2r R2HC/HC2R semantics), Pres_2::solve (src/pres_2.cxx:266-362), tdma (:202-263).
//
// Spectral layout (ours, internal): S[k][j][m] complex (interleaved re,im), m = 0..itot/2
// (nm = itot/2+1 x-modes); after the y transform j is the y-mode l = 0..jtot-1.
// The eigenvalue of slot (m,l) is bmati[m] + bmatj[l] exactly as in the reference's
// half-complex layout, so the tridiagonal systems are the reference's systems.
#pragma once
#include "common.cuh"

namespace mhh {

template <typename TF> struct cplx { TF x, y; };
template <typename TF> __device__ __forceinline__ cplx<TF> cmul(const cplx<TF> a, const cplx<TF> b)
{ return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
template <typename TF> __device__ __forceinline__ cplx<TF> cadd(const cplx<TF> a, const cplx<TF> b) { return {a.x + b.x, a.y + b.y}; }
template <typename TF> __device__ __forceinline__ cplx<TF> csub(const cplx<TF> a, const cplx<TF> b) { return {a.x - b.x, a.y - b.y}; }
template <typename TF> __device__ __forceinline__ cplx<TF> cconj(const cplx<TF> a) { return {a.x, -a.y}; }
// multiply by -i (forward rotation)
template <typename TF> __device__ __forceinline__ cplx<TF> cmul_mi(const cplx<TF> a) { return {a.y, -a.x}; }

struct FftPlan
{
    int n;            // complex transform length
    int nstages;
    int radix[16];
    int log2s[16];    // log2 of the stride entering stage st, or -1 when it is not a power of two
    int log2nb[16];   // log2 of the butterflies per sequence (n / radix), or -1
};

// ------------------------------------------------------------------------------------------
// Spectral workspace layout, single GPU and y-slab decomposition (npx = 1, npy = P).
//   x side (after the x transform): rank-local rows r = jl + k*jmax, all nm x-modes, stored as P
//     blocks -- block d holds the modes owned by rank d, [r][m - m_off(d)] -- so that block d IS the
//     all-to-all message for rank d (no pack pass).
//   y side (after the all-to-all): this rank's mcl modes for ALL jtot rows, stored as P blocks --
//     block s is the message received from rank s, [k*jmax + jl][ml] -- so the y transform and the
//     tridiagonal solve read the receive buffer directly (no unpack pass).
//   Modes are dealt out evenly: base = nm / P each, the first nm % P ranks get one more.
//   P = 1 degenerates to S[k][j][m] with pitch nm (one buffer, no exchange).
// Semantic model: reference src/transpose.cxx:117-271 (exec_xy / exec_yx) + src/fft.cxx:455-587.
// ------------------------------------------------------------------------------------------
struct SpecLayout
{
    int P;        // ranks along y
    int rank;
    int nm;       // x-modes in total (itot/2 + 1)
    int base, rem;
    int jmax, jtot, ktot;
    int mcl;      // modes owned by this rank
    int m_off;    // first owned mode
    long long rows;   // jmax*ktot
    int xtiled;       // x side in 8-mode panels (fused peer transposes): block d = [k][panel][jl][8], so that the inverse y
                      // transform's panel stores are long contiguous runs in the row owner's memory (NVLink-friendly)

    __host__ __device__ int count(const int d) const { return base + (d < rem ? 1 : 0); }
    __host__ __device__ int offset(const int d) const { return d * base + (d < rem ? d : rem); }
    __host__ __device__ int owner(const int m) const
    {
        const int cut = rem * (base + 1);
        return m < cut ? m / (base + 1) : rem + (m - cut) / base;
    }
    __host__ __device__ int pcnt(const int d) const { return (count(d) + 7) >> 3; }       // 8-mode panels of block d
    __host__ __device__ long long xoff_tiled(const int d) const
    {
        const int lo = (base + 7) >> 3, hi = (base + 8) >> 3;
        return 8 * rows * ((long long)d * lo + (long long)(d < rem ? d : rem) * (hi - lo));
    }
    // x side: complex index of (level k, local row jl, mode m)
    __host__ __device__ long long xidx_kj(const int k, const int jl, const int m) const
    {
        const long long r = (long long)k * jmax + jl;
        if (P == 1) return r * nm + m;
        const int d = owner(m);
        const int o = offset(d);
        if (!xtiled) return (long long)o * rows + r * count(d) + (m - o);
        const int ml = m - o;
        return xoff_tiled(d) + (((long long)k * pcnt(d) + (ml >> 3)) * jmax + jl) * 8 + (ml & 7);
    }
    __host__ __device__ long long xidx(const long long r, const int m) const
    {
        const int k = (int)(r / jmax);
        return xidx_kj(k, (int)(r - (long long)k * jmax), m);
    }
    __host__ __device__ long long xside_elems() const { return xtiled ? xoff_tiled(P) : (long long)nm * rows; }
    // y side: complex index of (level k, global row j, local mode ml)
    __host__ __device__ long long yidx(const int k, const int j, const int ml) const
    {
        if (P == 1) return ((long long)k * jtot + j) * nm + ml;
        const int s = j / jmax;
        const int jl = j - s * jmax;
        return (long long)s * mcl * rows + ((long long)k * jmax + jl) * mcl + ml;
    }
    // y side: a column (fixed j, ml) advances by this many complex elements per level
    __host__ __device__ long long ykstride() const { return (long long)jmax * mcl; }
};

inline SpecLayout make_spec_layout(int itot, int jtot, int ktot, int P, int rank)
{
    SpecLayout s{};
    s.P = P; s.rank = rank; s.nm = itot / 2 + 1;
    s.base = s.nm / P; s.rem = s.nm % P;
    s.jmax = jtot / P; s.jtot = jtot; s.ktot = ktot;
    s.mcl = s.count(rank); s.m_off = s.offset(rank);
    s.rows = (long long)s.jmax * ktot;
    return s;
}

// Peer-mapped spectral workspaces of all slab ranks (CUDA IPC over NVLink/NVSwitch).  When `on`, the x transform stores
// its modes straight into the y-side buffer of the rank that owns them and the inverse y transform stores straight into
// the x-side buffer of the rank that owns the row: the all-to-all IS the store phase of the FFT kernels (no separate
// transpose pass, no staging), followed by one tiny all-reduce as a barrier.
constexpr int MAX_SLAB_RANKS = 16;
template <typename TF>
struct PeerPtrs
{
    int on;
    TF* x[MAX_SLAB_RANKS];      // x-side buffers (complex, as TF pairs)
    TF* y[MAX_SLAB_RANKS];      // y-side buffers
};

// complex element address of (local row r, mode m) in the y-side buffer of the mode's owner
template <typename TF>
__device__ __forceinline__ cplx<TF>* peer_y_slot(const PeerPtrs<TF>& pp, const SpecLayout& lay, const long long r, const int m)
{
    const int d = lay.owner(m);
    const int cnt = lay.count(d);
    return reinterpret_cast<cplx<TF>*>(pp.y[d]) + ((long long)lay.rank * cnt + 0) * lay.rows + r * cnt + (m - lay.offset(d));
}

// complex element address of (level k, global row j, local mode ml) in the x-side buffer of the row's owner
template <typename TF>
__device__ __forceinline__ cplx<TF>* peer_x_slot(const PeerPtrs<TF>& pp, const SpecLayout& lay, const int k, const int j, const int ml)
{
    const int s = j / lay.jmax;
    const int jl = j - s * lay.jmax;
    // block `rank` of the (tiled) x side of rank s: [k][panel][jl][8]
    return reinterpret_cast<cplx<TF>*>(pp.x[s]) + lay.xoff_tiled(lay.rank)
           + (((long long)k * lay.pcnt(lay.rank) + (ml >> 3)) * lay.jmax + jl) * 8 + (ml & 7);
}

// integer division / remainder by a value whose log2 is known (or -1 -> generic)
__device__ __forceinline__ void divmod(const int x, const int d, const int lg, int& quo, int& rem)
{
    if (lg >= 0) { quo = x >> lg; rem = x & (d - 1); }
    else { quo = x / d; rem = x - quo * d; }
}

// One radix-R Stockham (decimation in frequency, autosort) butterfly.
//   x: input (length n), y: output.  s = stride (product of previous radices), m = n_cur / R.
//   tw: table of exp(-2 pi i t / n), t = 0..n-1 (full-length roots of unity).
template <typename TF, int R>
__device__ __forceinline__ void butterfly(const cplx<TF>* __restrict__ x, cplx<TF>* __restrict__ y,
        const cplx<TF>* __restrict__ tw, const int n, const int s, const int m, const int p, const int q)
{
    cplx<TF> a[R];
#pragma unroll
    for (int r = 0; r < R; ++r) a[r] = x[q + s * (p + r * m)];

    cplx<TF> b[R];
    if (R == 2)
    {
        b[0] = cadd(a[0], a[1]);
        b[1] = csub(a[0], a[1]);
    }
    else if (R == 4)
    {
        const cplx<TF> t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]);
        const cplx<TF> t2 = cadd(a[1], a[3]), t3 = cmul_mi(csub(a[1], a[3]));
        b[0] = cadd(t0, t2); b[2] = csub(t0, t2);
        b[1] = cadd(t1, t3); b[3] = csub(t1, t3);
    }
    else if (R == 8)
    {
        const TF h = TF(0.70710678118654752440);
        // radix-2 x radix-4 decomposition
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
        // twiddles w8^k: 1, (1-i)/sqrt2, -i, (-1-i)/sqrt2
        const cplx<TF> o1 = {h * (o[1].x + o[1].y), h * (o[1].y - o[1].x)};
        const cplx<TF> o2 = cmul_mi(o[2]);
        const cplx<TF> o3 = {h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y)};
        b[0] = cadd(e[0], o[0]); b[4] = csub(e[0], o[0]);
        b[1] = cadd(e[1], o1);   b[5] = csub(e[1], o1);
        b[2] = cadd(e[2], o2);   b[6] = csub(e[2], o2);
        b[3] = cadd(e[3], o3);   b[7] = csub(e[3], o3);
    }
    else if (R == 3)
    {
        const TF c = TF(-0.5), sn = TF(0.86602540378443864676);
        const cplx<TF> t1 = cadd(a[1], a[2]);
        const cplx<TF> t2 = {a[0].x + c * t1.x, a[0].y + c * t1.y};
        const cplx<TF> d = csub(a[1], a[2]);
        const cplx<TF> t3 = {sn * d.y, -sn * d.x};       // -i*sn*d
        b[0] = cadd(a[0], t1);
        b[1] = cadd(t2, t3);
        b[2] = csub(t2, t3);
    }
    else if (R == 5)
    {
        const TF c1 = TF(0.30901699437494742410), c2 = TF(-0.80901699437494742410);
        const TF s1 = TF(0.95105651629515357212), s2 = TF(0.58778525229247312917);
        const cplx<TF> t1 = cadd(a[1], a[4]), t2 = cadd(a[2], a[3]);
        const cplx<TF> d1 = csub(a[1], a[4]), d2 = csub(a[2], a[3]);
        b[0] = {a[0].x + t1.x + t2.x, a[0].y + t1.y + t2.y};
        const cplx<TF> m1 = {a[0].x + c1 * t1.x + c2 * t2.x, a[0].y + c1 * t1.y + c2 * t2.y};
        const cplx<TF> m2 = {a[0].x + c2 * t1.x + c1 * t2.x, a[0].y + c2 * t1.y + c1 * t2.y};
        // -i*(s1*d1 + s2*d2), -i*(s2*d1 - s1*d2)
        const cplx<TF> n1 = {s1 * d1.y + s2 * d2.y, -(s1 * d1.x + s2 * d2.x)};
        const cplx<TF> n2 = {s2 * d1.y - s1 * d2.y, -(s2 * d1.x - s1 * d2.x)};
        b[1] = cadd(m1, n1); b[4] = csub(m1, n1);
        b[2] = cadd(m2, n2); b[3] = csub(m2, n2);
    }

    // twiddle: w_{n_cur}^{p*r} = tw[p*r*s]   (p < m = n_cur/R and n_cur*s = n, so p*r*s < n)
    cplx<TF>* yo = y + q + s * (R * p);
    yo[0] = b[0];
    if (p == 0)
    {
#pragma unroll
        for (int r = 1; r < R; ++r) yo[s * r] = b[r];
    }
    else
    {
        const int ps = p * s;
#pragma unroll
        for (int r = 1; r < R; ++r) yo[s * r] = cmul(b[r], tw[ps * r]);
    }
}

// Complex forward FFT of `nseq` sequences of length plan.n held in shared memory.
// seq stride = ld (complex elements).  Ping-pongs between buf0 and buf1; returns the buffer
// holding the result (natural order).  All threads of the CTA must call.
template <typename TF>
__device__ cplx<TF>* smem_fft(cplx<TF>* buf0, cplx<TF>* buf1, const cplx<TF>* __restrict__ tw,
        const FftPlan& plan, const int nseq, const int ld)
{
    const int n = plan.n;
    const int tid = threadIdx.x;
    const int nth = blockDim.x;
    int s = 1;
    int ncur = n;
    cplx<TF>* x = buf0;
    cplx<TF>* y = buf1;
    for (int st = 0; st < plan.nstages; ++st)
    {
        const int R = plan.radix[st];
        const int m = ncur / R;
        const int nb = n / R;              // butterflies per sequence
        const int lgs = plan.log2s[st], lgnb = plan.log2nb[st];
        for (int t = tid; t < nseq * nb; t += nth)
        {
            int sq, b, p, q;
            divmod(t, nb, lgnb, sq, b);
            divmod(b, s, lgs, p, q);
            const cplx<TF>* xs = x + sq * ld;
            cplx<TF>* ys = y + sq * ld;
            switch (R)
            {
                case 8: butterfly<TF, 8>(xs, ys, tw, n, s, m, p, q); break;
                case 4: butterfly<TF, 4>(xs, ys, tw, n, s, m, p, q); break;
                case 2: butterfly<TF, 2>(xs, ys, tw, n, s, m, p, q); break;
                case 3: butterfly<TF, 3>(xs, ys, tw, n, s, m, p, q); break;
                default: butterfly<TF, 5>(xs, ys, tw, n, s, m, p, q); break;
            }
        }
        __syncthreads();
        cplx<TF>* tmp = x; x = y; y = tmp;
        s *= R;
        ncur = m;
    }
    return x;
}

// ------------------------------------------------------------------------------------------
// x transform, forward: rows of itot reals -> nm = itot/2+1 complex modes (in place in `spec`,
// row pitch 2*nm reals).  Real FFT through a half-length complex FFT.
// RHS_FUSED: the row is not read from `spec` but computed on the fly as the pressure rhs
// (Pres_2::input), so the divergence never makes a round trip through HBM.
// One CTA handles ROWS rows at a time; grid-strides over all jtot*ktot rows.
// ------------------------------------------------------------------------------------------
template <typename TF>
struct RhsSrc
{
    const TF* u; const TF* v; const TF* w; const TF* ut; const TF* vt; const TF* wt;
    TF dti;
    int ywrap;      // 1: the block is periodic in y by itself (single GPU); 0: read the exchanged north ghost row of vt / v
};

template <typename TF, bool RHS_FUSED>
__global__ void fft_x_forward_kernel(TF* __restrict__ spec, const RhsSrc<TF> src, const GridDev<TF> g, const SpecLayout lay, const PeerPtrs<TF> pp,
        const FftPlan plan, const cplx<TF>* __restrict__ tw_half, const cplx<TF>* __restrict__ tw_full,
        const int rows_per_cta, const long long nrows)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = plan.n;               // itot/2
    const int N = 2 * L;
    const int nm = L + 1;
    const int ld = L + 1;               // padded sequence stride
    cplx<TF>* buf0 = reinterpret_cast<cplx<TF>*>(smem_raw);
    cplx<TF>* buf1 = buf0 + rows_per_cta * ld;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const long long jj = g.icells, kk = g.ijcells;

    for (long long row0 = (long long)blockIdx.x * rows_per_cta; row0 < nrows; row0 += (long long)gridDim.x * rows_per_cta)
    {
        const int nr = (int)min((long long)rows_per_cta, nrows - row0);
        // load: z[n] = x[2n] + i x[2n+1]; one warp per row so that the row decode is done once
        const int nchN = (N + 31) >> 5;
        for (int wi = warp; wi < nr * nchN; wi += nwarps)
        {
            const int r = wi / nchN;
            const int i0 = (wi - r * nchN) << 5;
            const long long row = row0 + r;
            TF* dst = reinterpret_cast<TF*>(buf0 + r * ld);
            if (RHS_FUSED)
            {
                const int kq = (int)(row / g.jmax);
                const int k = kq + g.kstart;
                const int j = (int)(row - (long long)kq * g.jmax) + g.jstart;
                const long long base = g.istart + j * jj + k * kk;
                const long long jn_off = (src.ywrap && j + 1 == g.jend) ? (1 - g.jmax) * jj : jj;
                const TF dti = src.dti;
                const TF rho = g.rhoref[k], rhoh0 = g.rhorefh[k], rhoh1 = g.rhorefh[k + 1], dzi = g.dzi[k];
                const int i = i0 + lane;
                if (i < N)
                {
                    const long long ijk = base + i;
                    const long long ie = (i + 1 == N) ? ijk + 1 - N : ijk + 1;
                    const long long jn = ijk + jn_off;
                    dst[i] = rho * ((src.ut[ie] + src.u[ie] * dti) - (src.ut[ijk] + src.u[ijk] * dti)) * g.dxi
                           + rho * ((src.vt[jn] + src.v[jn] * dti) - (src.vt[ijk] + src.v[ijk] * dti)) * g.dyi
                           + (rhoh1 * (src.wt[ijk + kk] + src.w[ijk + kk] * dti)
                            - rhoh0 * (src.wt[ijk     ] + src.w[ijk     ] * dti)) * dzi;
                }
            }
            else
            {
                const TF* in = src.u ? src.u + row * (long long)N : spec + row * (2 * nm);     // separate compact source when src.u is set
                const int i = i0 + lane;
                if (i < N) dst[i] = in[i];
            }
        }
        __syncthreads();
        cplx<TF>* Z = smem_fft<TF>(buf0, buf1, tw_half, plan, nr, ld);
        // post-process: X[m] = E[m] + W_N^m O[m]
        const int nchM = (nm + 31) >> 5;
        for (int wi = warp; wi < nr * nchM; wi += nwarps)
        {
            const int r = wi / nchM;
            const int m = ((wi - r * nchM) << 5) + lane;
            const cplx<TF>* z = Z + r * ld;
            cplx<TF>* out = reinterpret_cast<cplx<TF>*>(spec);
            if (m < nm)
            {
                const cplx<TF> zm = z[m == L ? 0 : m];
                const cplx<TF> zc = cconj(z[m == 0 ? 0 : L - m]);
                const cplx<TF> e = {TF(0.5) * (zm.x + zc.x), TF(0.5) * (zm.y + zc.y)};
                const cplx<TF> d = {TF(0.5) * (zm.x - zc.x), TF(0.5) * (zm.y - zc.y)};
                const cplx<TF> o = cmul_mi(d);                  // (zm - zc)/(2i)
                const cplx<TF> X = cadd(e, cmul(tw_full[m], o));           // tw_full[m] = exp(-2 pi i m / N)
                if (pp.on) *peer_y_slot<TF>(pp, lay, row0 + r, m) = X; else out[lay.xidx(row0 + r, m)] = X;
            }
        }
        __syncthreads();
    }
    if (pp.on) __threadfence_system();
}

// ------------------------------------------------------------------------------------------
// x transform, backward, fused with Pres_2::solve's unpack: half-spectrum rows -> real rows
// scaled by `norm` = 1/(itot*jtot), written straight into the ghosted pressure array including
// the periodic ghost cells in x and y and the zero-gradient bottom ghost level
// (src/pres_2.cxx:339-361), so no separate copy / boundary_cyclic pass over p is needed.
// ------------------------------------------------------------------------------------------
template <typename TF>
__global__ void fft_x_backward_kernel(const TF* __restrict__ spec, TF* __restrict__ p, const GridDev<TF> g, const SpecLayout lay,
        const FftPlan plan, const cplx<TF>* __restrict__ tw_half, const cplx<TF>* __restrict__ tw_full,
        const int rows_per_cta, const long long nrows, const TF norm, const int fill_y_ghosts)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = plan.n;
    const int ld = L + 1;
    cplx<TF>* buf0 = reinterpret_cast<cplx<TF>*>(smem_raw);
    cplx<TF>* buf1 = buf0 + rows_per_cta * ld;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const long long jj = g.icells, kk = g.ijcells;

    for (long long row0 = (long long)blockIdx.x * rows_per_cta; row0 < nrows; row0 += (long long)gridDim.x * rows_per_cta)
    {
        const int nr = (int)min((long long)rows_per_cta, nrows - row0);
        // pre-process: Z'[m] = (X[m] + conj X[L-m]) + i e^{+2 pi i m/N} (X[m] - conj X[L-m]); store conj(Z') so the
        // forward transform yields conj(inverse).
        const int nchL = (L + 31) >> 5;
        for (int wi = warp; wi < nr * nchL; wi += nwarps)
        {
            const int r = wi / nchL;
            const int m = ((wi - r * nchL) << 5) + lane;
            const cplx<TF>* X = reinterpret_cast<const cplx<TF>*>(spec);
            if (m < L)
            {
                const cplx<TF> xm = X[lay.xidx(row0 + r, m)];
                const cplx<TF> xc = cconj(X[lay.xidx(row0 + r, L - m)]);
                const cplx<TF> e = cadd(xm, xc);
                const cplx<TF> d = csub(xm, xc);
                const cplx<TF> wd = cmul(cconj(tw_full[m]), d);  // exp(+2 pi i m / N) * d
                buf0[r * ld + m] = {e.x - wd.y, -(e.y + wd.x)};   // conj(e + i*wd)
            }
        }
        __syncthreads();
        cplx<TF>* Z = smem_fft<TF>(buf0, buf1, tw_half, plan, nr, ld);
        // store: x[2n] = Re(conj Z[n]) = Z[n].x, x[2n+1] = Im(conj Z[n]) = -Z[n].y
        const int wtot = g.itot + 2 * g.igc;
        const int nchW = (wtot + 31) >> 5;
        for (int wi = warp; wi < nr * nchW; wi += nwarps)
        {
            const int r = wi / nchW;
            const int ic = ((wi - r * nchW) << 5) + lane;
            const long long row = row0 + r;
            const int kq = (int)(row / g.jmax);
            const int jq = (int)(row - (long long)kq * g.jmax);
            const long long rowbase = (jq + g.jstart) * jj + (kq + g.kstart) * kk;
            const bool ylo = fill_y_ghosts && jq < g.jgc, yhi = fill_y_ghosts && jq >= g.jmax - g.jgc;
            const cplx<TF>* z = Z + r * ld;
            if (ic < wtot)
            {
                int i = ic - g.igc;                   // interior index, wrapped
                if (i < 0) i += g.itot; else if (i >= g.itot) i -= g.itot;
                const cplx<TF> zz = z[i >> 1];
                const TF val = ((i & 1) ? -zz.y : zz.x) * norm;
                const long long base = ic + rowbase;
                p[base] = val;
                if (kq == 0) p[base - kk] = val;
                if (ylo)
                {
                    p[base + g.jmax * jj] = val;
                    if (kq == 0) p[base + g.jmax * jj - kk] = val;
                }
                if (yhi)
                {
                    p[base - g.jmax * jj] = val;
                    if (kq == 0) p[base - g.jmax * jj - kk] = val;
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// y transform (complex, length jtot, stride = row pitch) on panels of MC consecutive x-modes.
// inverse != 0 -> unnormalised inverse (conjugate trick).
// ------------------------------------------------------------------------------------------
template <typename TF>
__global__ void fft_y_kernel(TF* __restrict__ spec, const SpecLayout lay, const PeerPtrs<TF> pp, const int nm, const int jtot, const int ktot,
        const FftPlan plan, const cplx<TF>* __restrict__ tw, const int MC, const int inverse)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ld = jtot + 1;
    cplx<TF>* buf0 = reinterpret_cast<cplx<TF>*>(smem_raw);
    cplx<TF>* buf1 = buf0 + MC * ld;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int npanel_m = (nm + MC - 1) / MC;
    const int lgMC = 31 - __clz(MC);          // MC is a power of two
    const long long npanels = (long long)npanel_m * ktot;
    cplx<TF>* S = reinterpret_cast<cplx<TF>*>(spec);

    for (long long pnl = blockIdx.x; pnl < npanels; pnl += gridDim.x)
    {
        const int k = (int)(pnl / npanel_m);
        const int m0 = (int)(pnl % npanel_m) * MC;
        const int mc = min(MC, nm - m0);
        for (int t = tid; t < jtot * MC; t += nth)
        {
            const int j = t >> lgMC;
            const int c = t & (MC - 1);
            if (c < mc)
            {
                cplx<TF> v = S[lay.yidx(k, j, m0 + c)];
                if (inverse) v.y = -v.y;
                buf0[c * ld + j] = v;
            }
        }
        __syncthreads();
        cplx<TF>* Z = smem_fft<TF>(buf0, buf1, tw, plan, mc, ld);
        for (int t = tid; t < jtot * MC; t += nth)
        {
            const int j = t >> lgMC;
            const int c = t & (MC - 1);
            if (c < mc)
            {
                cplx<TF> v = Z[c * ld + j];
                if (inverse) v.y = -v.y;
                if (pp.on && inverse) *peer_x_slot<TF>(pp, lay, k, j, m0 + c) = v; else S[lay.yidx(k, j, m0 + c)] = v;
            }
        }
        __syncthreads();
    }
    if (pp.on && inverse) __threadfence_system();
}

// ------------------------------------------------------------------------------------------
// Tridiagonal solve in z, one thread per (l, m) complex column (re and im share the factors).
// The elimination factors depend only on the grid and the base state, so they are tabulated
// once (`fac`, [k][l][m]) by tdma_setup_kernel and the sweeps carry no division chain.
//   reference tdma:  work3d[k] = c[k-1]/w[k-1];  w[k] = b[k] - a[k]*work3d[k];
//                    p[k] = (p[k] - a[k]*p[k-1]) / w[k];   back: p[k] -= work3d[k+1]*p[k+1]
//   table: fac[k] = work3d[k] (k>=1), fac[0] unused;  winv[k] = 1/w[k] is rebuilt from b and fac.
// ------------------------------------------------------------------------------------------
template <typename TF>
struct TdmaCoef
{
    const TF* a; const TF* c;        // kmax
    const TF* dz2rho;                // kmax: dz^2 * rhoref
    const TF* dz2;                   // kmax: dz^2
    const TF* bmati; const TF* bmatj; // nm, jtot
};

template <typename TF>
__device__ __forceinline__ TF tdma_b(const TdmaCoef<TF>& cf, const int k, const int kmax, const TF lam, const bool mode00)
{
    TF b = cf.dz2rho[k] * lam - (cf.a[k] + cf.c[k]);
    if (k == 0) b += cf.a[0];
    if (k == kmax - 1) b = mode00 ? b - cf.c[kmax - 1] : b + cf.c[kmax - 1];
    return b;
}

template <typename TF>
__global__ void tdma_setup_kernel(TF* __restrict__ fac, const TdmaCoef<TF> cf, const int nm, const int jtot, const int kmax,
        const int m_off, const int l_off)
{
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long ncol = (long long)nm * jtot;
    if (col >= ncol) return;
    const int l = (int)(col / nm), m = (int)(col % nm);
    const TF lam = cf.bmati[m + m_off] + cf.bmatj[l + l_off];
    const bool mode00 = (m + m_off == 0) && (l + l_off == 0);
    TF w = tdma_b(cf, 0, kmax, lam, mode00);
    fac[col] = TF(0);
    for (int k = 1; k < kmax; ++k)
    {
        const TF f = cf.c[k - 1] / w;
        fac[col + k * ncol] = f;
        w = tdma_b(cf, k, kmax, lam, mode00) - cf.a[k] * f;
    }
}

template <typename TF>
__global__ void __launch_bounds__(128) tdma_solve_kernel(TF* __restrict__ spec, const TF* __restrict__ fac, const TdmaCoef<TF> cf,
        const SpecLayout lay, const int nm, const int jtot, const int kmax, const int m_off, const int l_off)
{
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long ncol = (long long)nm * jtot;
    if (col >= ncol) return;
    const int l = (int)(col / nm), m = (int)(col % nm);
    const TF lam = cf.bmati[m + m_off] + cf.bmatj[l + l_off];
    const bool mode00 = (m + m_off == 0) && (l + l_off == 0);
    cplx<TF>* S = reinterpret_cast<cplx<TF>*>(spec) + lay.yidx(0, l, m);
    const long long ks = lay.ykstride();

    // forward sweep
    cplx<TF> prev;
    {
        const TF w = tdma_b(cf, 0, kmax, lam, mode00);
        cplx<TF> v = S[0];
        const TF s = cf.dz2[0] / w;
        prev = {v.x * s, v.y * s};
        S[0] = prev;
    }
#pragma unroll 4
    for (int k = 1; k < kmax; ++k)
    {
        const TF f = fac[col + k * ncol];
        const TF ak = cf.a[k];
        const TF w = tdma_b(cf, k, kmax, lam, mode00) - ak * f;
        const TF winv = TF(1) / w;
        const cplx<TF> v = S[k * ks];
        const TF d2 = cf.dz2[k];
        prev = {(d2 * v.x - ak * prev.x) * winv, (d2 * v.y - ak * prev.y) * winv};
        S[k * ks] = prev;
    }
    // back substitution
#pragma unroll 4
    for (int k = kmax - 2; k >= 0; --k)
    {
        const TF f = fac[col + (k + 1) * ncol];
        const cplx<TF> v = S[k * ks];
        prev = {v.x - f * prev.x, v.y - f * prev.y};
        S[k * ks] = prev;
    }
}

} // namespace mhh
