// mhhb200 -- host drivers of the pressure solvers (Pres_2 three-kernel and fused versions, Pres_4).
#include "host_common.cuh"

namespace mhhhost {

template <typename TF>
int twiddles(mhh_ctx* c, cplx<TF>** out, int n)
{
    std::vector<cplx<TF>> h((size_t)std::max(n, 1));
    for (int t = 0; t < n; ++t)
    {
        // exact octant symmetries keep the table accurate to the last bit
        const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)t / (long double)n;
        h[t].x = (TF)cosl(ang);
        h[t].y = (TF)sinl(ang);
    }
    CUDA_TRY(c, cudaMalloc(out, sizeof(cplx<TF>) * std::max(n, 1)));
    CUDA_TRY(c, cudaMemcpy(*out, h.data(), sizeof(cplx<TF>) * std::max(n, 1), cudaMemcpyHostToDevice));
    return MHH_OK;
}

// ---- warp-per-sequence FFT dispatch (fft_warp.cuh) -------------------------------------------
#define WFFT_X_CASES(X) X(16) X(32) X(64) X(128) X(256) X(512) X(1024)
#define WFFT_Y_CASES(X) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)

template <typename TF>
int wfft_x_attrs(mhh_ctx* c, int L)
{
    switch (L)
    {
#define X(N) case N: \
        CUDA_TRY(c, cudaFuncSetAttribute(wfft_x_forward_kernel<TF, N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wfft_smem<TF, N>())); \
        CUDA_TRY(c, cudaFuncSetAttribute(wfft_x_forward_kernel<TF, N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wfft_smem<TF, N>())); \
        CUDA_TRY(c, cudaFuncSetAttribute(wfft_x_backward_kernel<TF, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wfft_smem<TF, N>())); break;
        WFFT_X_CASES(X)
#undef X
        default: c->err = "wfft_x: unsupported length"; return MHH_E_INVALID;
    }
    return MHH_OK;
}

template <typename TF>
int wfft_y_attrs(mhh_ctx* c, int J)
{
    switch (J)
    {
#define X(N) case N: CUDA_TRY(c, cudaFuncSetAttribute(wfft_y_kernel<TF, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wfft_smem<TF, N>())); break;
        WFFT_Y_CASES(X)
#undef X
        default: c->err = "wfft_y: unsupported length"; return MHH_E_INVALID;
    }
    return MHH_OK;
}

template <typename TF>
void wfft_x_forward_launch(int L, bool fused, int grid, cudaStream_t st, TF* spec, const RhsSrc<TF>& src, const GridDev<TF>& g, const SpecLayout& lay, const PeerPtrs<TF>& pp,
                           const cplx<TF>* twh, const cplx<TF>* twf, long long nrows)
{
    switch (L)
    {
#define X(N) case N: if (fused) wfft_x_forward_kernel<TF, N, true><<<grid, 32 * WFFT_WARPS, wfft_smem<TF, N>(), st>>>(spec, src, g, lay, pp, twh, twf, nrows); \
                     else wfft_x_forward_kernel<TF, N, false><<<grid, 32 * WFFT_WARPS, wfft_smem<TF, N>(), st>>>(spec, src, g, lay, pp, twh, twf, nrows); break;
        WFFT_X_CASES(X)
#undef X
    }
}

template <typename TF>
void wfft_x_backward_launch(int L, int grid, cudaStream_t st, const TF* spec, TF* p, const GridDev<TF>& g, const SpecLayout& lay,
                            const cplx<TF>* twh, const cplx<TF>* twf, long long nrows, TF norm, int fill)
{
    switch (L)
    {
#define X(N) case N: wfft_x_backward_kernel<TF, N><<<grid, 32 * WFFT_WARPS, wfft_smem<TF, N>(), st>>>(spec, p, g, lay, twh, twf, nrows, norm, fill); break;
        WFFT_X_CASES(X)
#undef X
    }
}

template <typename TF>
void wfft_y_launch(int J, int grid, cudaStream_t st, TF* spec, const SpecLayout& lay, const PeerPtrs<TF>& pp, int nm, int ktot, const cplx<TF>* tw, int inverse)
{
    switch (J)
    {
#define X(N) case N: wfft_y_kernel<TF, N><<<grid, 32 * WFFT_WARPS, wfft_smem<TF, N>(), st>>>(spec, lay, pp, nm, ktot, tw, inverse); break;
        WFFT_Y_CASES(X)
#undef X
    }
}

// ---- Pres_2 version 2 dispatch (poisson_fused.cuh) ----------------------------------------------
template <typename TF>
int p2_attrs(mhh_ctx* c, int L, int J)
{
    switch (L)
    {
#define X(N) case N: \
        CUDA_TRY(c, cudaFuncSetAttribute(p2_x_forward_kernel<TF, N, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wfft_smem<TF, N>())); \
        CUDA_TRY(c, cudaFuncSetAttribute(p2_x_forward_kernel<TF, N, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wfft_smem<TF, N>())); \
        CUDA_TRY(c, cudaFuncSetAttribute(p2_x_forward_kernel<TF, N, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wfft_smem<TF, N>())); \
        CUDA_TRY(c, cudaFuncSetAttribute(p2_x_backward_kernel<TF, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wfft_smem<TF, N>())); break;
        WFFT_X_CASES(X)
#undef X
        default: c->err = "p2: unsupported x length"; return MHH_E_INVALID;
    }
    switch (J)
    {
#define X(N) case N: \
        CUDA_TRY(c, cudaFuncSetAttribute(p2_y_forward_kernel<TF, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p2_y_smem<TF, N>())); \
        CUDA_TRY(c, cudaFuncSetAttribute(p2_y_backward_kernel<TF, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p2_y_smem<TF, N>())); break;
        WFFT_Y_CASES(X)
#undef X
        default: c->err = "p2: unsupported y length"; return MHH_E_INVALID;
    }
    return MHH_OK;
}

template <typename TF>
size_t p2_y_smem_of(int J)
{
    switch (J)
    {
#define X(N) case N: return p2_y_smem<TF, N>();
        WFFT_Y_CASES(X)
#undef X
    }
    return (size_t)1 << 30;
}

template <typename TF>
void p2_x_forward_launch(int L, bool fused_rhs, int grid, cudaStream_t st, const TF* compact, const RhsSrc<TF>& src, const GridDev<TF>& g,
                         const Spec2& lay, const XferPtrs<TF>& xf, const cplx<TF>* twh, const cplx<TF>* twf)
{
    // vector loads of the x-pairs: every pair must start on a 2*sizeof(TF) boundary (MHH_RHS_VEC=0 keeps the scalar loader)
    static const bool allow_vec = [] { const char* e = getenv("MHH_RHS_VEC"); return !(e && atoi(e) == 0); }();
    // measured at 512^3: fp32 3.97 -> 3.80 ms/step, fp64 5.55 -> 5.71 (16-byte loads at a 32-byte lane stride): fp32 only
    bool vec = allow_vec && sizeof(TF) == 4 && fused_rhs && (g.istart % 2 == 0) && (g.icells % 2 == 0);
    if (vec)
        for (const void* p : {(const void*)src.u, (const void*)src.v, (const void*)src.w, (const void*)src.ut, (const void*)src.vt, (const void*)src.wt})
            if (reinterpret_cast<uintptr_t>(p) % (2 * sizeof(TF)) != 0) vec = false;
    switch (L)
    {
#define X(N) case N: if (fused_rhs && vec) p2_x_forward_kernel<TF, N, 2><<<grid, 32 * P2_ROWS, wfft_smem<TF, N>(), st>>>(compact, src, g, lay, xf, twh, twf); \
                     else if (fused_rhs) p2_x_forward_kernel<TF, N, 1><<<grid, 32 * P2_ROWS, wfft_smem<TF, N>(), st>>>(compact, src, g, lay, xf, twh, twf); \
                     else p2_x_forward_kernel<TF, N, 0><<<grid, 32 * P2_ROWS, wfft_smem<TF, N>(), st>>>(compact, src, g, lay, xf, twh, twf); break;
        WFFT_X_CASES(X)
#undef X
    }
}

template <typename TF>
void p2_x_backward_launch(int L, int grid, cudaStream_t st, const cplx<TF>* X_, TF* p, const GridDev<TF>& g, const Spec2& lay,
                          const cplx<TF>* twh, const cplx<TF>* twf, TF norm, int fill)
{
    switch (L)
    {
#define X(N) case N: p2_x_backward_kernel<TF, N><<<grid, 32 * P2_ROWS, wfft_smem<TF, N>(), st>>>(X_, p, g, lay, twh, twf, norm, fill); break;
        WFFT_X_CASES(X)
#undef X
    }
}

template <typename TF>
void p2_y_launch(int J, bool backward, cudaStream_t st, cplx<TF>* Y, const TF* T, const Spec2& lay, const XferPtrs<TF>& xf,
                 const TdmaCoef<TF>& cf, const cplx<TF>* tw, int jlog2, int solve)
{
    switch (J)
    {
#define X(N) case N: if (backward) p2_y_backward_kernel<TF, N><<<2 * lay.mcl, 32 * p2_y_warps<TF, N>(), p2_y_smem<TF, N>(), st>>>(Y, T, T + (size_t)lay.mcl * lay.ktot * N, lay, xf, cf.a, cf.c, tw, jlog2, solve); \
                     else p2_y_forward_kernel<TF, N><<<2 * lay.mcl, 32 * p2_y_warps<TF, N>(), p2_y_smem<TF, N>(), st>>>(Y, T, lay, cf.a, cf.c, cf.dz2, tw, jlog2, solve); break;
        WFFT_Y_CASES(X)
#undef X
    }
}

template <typename TF>
void p2_setup_launch(int J, cudaStream_t st, TF* T, const TdmaCoef<TF>& cf, int mcl, int kmax, int m_off)
{
    const long long ncol = (long long)mcl * J;
    const unsigned grid = (unsigned)((ncol + 127) / 128);
    switch (J)
    {
#define X(N) case N: tdma2_setup_kernel<TF, N><<<grid, 128, 0, st>>>(T, T + (size_t)mcl * kmax * N, cf, mcl, kmax, m_off); break;
        WFFT_Y_CASES(X)
#undef X
    }
}

// Pres_2 plans, twiddles, workspace and launch geometry (called once from mhh_ctx_create)
template <typename TF>
int pres_create(Ctx<TF>* c)
{
    GridDev<TF>& g = c->g;
    const mhh_grid_desc* d = &c->desc;
    const int P = c->nranks;
    // ---- Pres_2 plans, twiddles, workspace ------------------------------------------------
    c->nm = g.itot / 2 + 1;
    bool okx = false, oky = false;
    c->plan_x = make_plan(g.itot / 2, okx);
    c->plan_y = make_plan(g.jtot, oky);
    if (!okx || !oky) { c->err = "itot/2 and jtot must factor into 2, 3 and 5"; return MHH_E_INVALID; }
    int rc;
    if ((rc = twiddles<TF>(c, &c->tw_xh, g.itot / 2)) != MHH_OK) return rc;
    if ((rc = twiddles<TF>(c, &c->tw_xf, g.itot)) != MHH_OK) return rc;
    if ((rc = twiddles<TF>(c, &c->tw_y, g.jtot)) != MHH_OK) return rc;

    // Pres_2 version 2 (y transforms fused with the Thomas sweeps): power-of-two itot/2 and jtot, 2nd-order grid.
    // MHH_PRES_FUSED=0 keeps the three-kernel version (A/B comparisons).
    {
        auto pow2_in = [](int v, int lo, int hi) { return v >= lo && v <= hi && (v & (v - 1)) == 0; };
        const bool off = (getenv("MHH_PRES_FUSED") && getenv("MHH_PRES_FUSED")[0] == '0') || (getenv("MHH_NO_WFFT") && getenv("MHH_NO_WFFT")[0] == '1');
        c->fused = !off && !g.dzi4 && pow2_in(g.itot / 2, 16, 1024) && pow2_in(g.jtot, 8, 2048)
                   && (size_t)WFFT_WARPS * (size_t)(fpad(g.itot / 2 - 1) + 2) * sizeof(cplx<TF>) <= (size_t)227 * 1024
                   && p2_y_smem_of<TF>(g.jtot) <= (size_t)227 * 1024;
        c->lay2 = make_spec2(g.itot, g.jtot, g.ktot, P, c->rank);
        c->jlog2 = 0; while ((1 << c->jlog2) < g.jmax) ++c->jlog2;
    }
    if (c->fused)
    {
        // pivot table T[ml][k][pos] followed by the interface factors Dinv[ml][pos]
        const size_t nX = (size_t)2 * c->lay2.xside_elems(), nY = (size_t)2 * c->lay2.yside_elems(), nT = (size_t)c->lay2.yside_elems() + (size_t)c->lay2.mcl * g.jtot;
        CUDA_TRY(c, cudaMalloc(&c->spec, sizeof(TF) * nX));
        CUDA_TRY(c, cudaMalloc(&c->specT, sizeof(TF) * nY));
        CUDA_TRY(c, cudaMalloc(&c->fac, sizeof(TF) * nT));
        c->ws_bytes = (long long)(sizeof(TF) * (nX + nY + nT));
    }
    else
    {
    // x side: room for the 8-mode-panel layout of the fused peer transposes (a few per cent of padding when P > 1)
    SpecLayout tiled = c->lay; tiled.xtiled = 1;
    const size_t nspec = (size_t)2 * std::max<long long>((long long)c->nm * g.jmax * g.ktot, P > 1 ? tiled.xside_elems() : 0);
    const size_t nspecT = (size_t)2 * c->lay.mcl * g.jtot * g.ktot;
    const size_t nfac = (size_t)c->lay.mcl * g.jtot * g.ktot;
    CUDA_TRY(c, cudaMalloc(&c->spec, sizeof(TF) * nspec));
    if (P > 1) CUDA_TRY(c, cudaMalloc(&c->specT, sizeof(TF) * nspecT)); else c->specT = c->spec;
    CUDA_TRY(c, cudaMalloc(&c->fac, sizeof(TF) * nfac));
    c->ws_bytes = (long long)(sizeof(TF) * (nspec + (P > 1 ? nspecT : 0) + nfac));
    }
    CUDA_TRY(c, cudaMalloc(&c->d_bmati, sizeof(TF) * c->nm));
    CUDA_TRY(c, cudaMalloc(&c->d_bmatj, sizeof(TF) * g.jtot));
    CUDA_TRY(c, cudaMalloc(&c->d_a, sizeof(TF) * g.kmax));
    CUDA_TRY(c, cudaMalloc(&c->d_c, sizeof(TF) * g.kmax));
    CUDA_TRY(c, cudaMalloc(&c->d_dz2rho, sizeof(TF) * g.kmax));
    CUDA_TRY(c, cudaMalloc(&c->d_dz2, sizeof(TF) * g.kmax));

    // launch geometry of the FFT kernels
    const int L = g.itot / 2;
    c->rows_x = std::max(1, std::min(64, 2048 / std::max(L, 1)));
    c->smem_x = (size_t)2 * c->rows_x * (L + 1) * sizeof(cplx<TF>);
    int mc = std::max(4, std::min(16, 2048 / g.jtot));
    while (mc & (mc - 1)) mc &= (mc - 1);      // power of two (the kernel uses shifts)
    if (sizeof(TF) == 4) mc *= 2;
    while (mc > 1 && (size_t)2 * mc * (g.jtot + 1) * sizeof(cplx<TF>) > 200 * 1024) mc /= 2;
    c->mc_y = mc;
    c->smem_y = (size_t)2 * mc * (g.jtot + 1) * sizeof(cplx<TF>);
    if (c->smem_x > 220 * 1024 || c->smem_y > 220 * 1024) { c->err = "grid too large for the shared-memory FFT"; return MHH_E_INVALID; }
    CUDA_TRY(c, cudaFuncSetAttribute(fft_x_forward_kernel<TF, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_x));
    CUDA_TRY(c, cudaFuncSetAttribute(fft_x_forward_kernel<TF, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_x));
    CUDA_TRY(c, cudaFuncSetAttribute(fft_x_backward_kernel<TF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_x));
    CUDA_TRY(c, cudaFuncSetAttribute(fft_y_kernel<TF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_y));
    // warp-per-sequence kernels for power-of-two lengths (MHH_NO_WFFT=1 keeps the generic block kernels)
    const bool no_wfft = getenv("MHH_NO_WFFT") && getenv("MHH_NO_WFFT")[0] == '1';
    auto pow2_in = [](int v, int lo, int hi) { return v >= lo && v <= hi && (v & (v - 1)) == 0; };
    // one warp-private padded row per warp must fit the 227 KB of shared memory (fp64: up to 1024 points, fp32: 2048)
    auto wfft_fits = [](int n) { return (size_t)WFFT_WARPS * (size_t)(fpad(n - 1) + 2) * sizeof(cplx<TF>) <= (size_t)227 * 1024; };
    c->wfft_x = !no_wfft && pow2_in(L, 16, 1024) && wfft_fits(L);
    c->wfft_y = !no_wfft && pow2_in(g.jtot, 8, 2048) && wfft_fits(g.jtot);
    int rc2;
    if (c->wfft_x && (rc2 = wfft_x_attrs<TF>(c, L)) != MHH_OK) return rc2;
    if (c->wfft_y && (rc2 = wfft_y_attrs<TF>(c, g.jtot)) != MHH_OK) return rc2;
    if (c->fused && (rc2 = p2_attrs<TF>(c, L, g.jtot)) != MHH_OK) return rc2;
    return MHH_OK;
}


// Pres_2::set_values (src/pres_2.cxx:124-153) and the per-mode pivot tables; called from mhh_set_basestate
template <typename TF>
int pres_set_values(Ctx<TF>* c)
{
    GridDev<TF>& g = c->g;
    const int kc = g.kcells; (void)kc;
    const TF* rr = c->h_rhoref.data(); const TF* rh = c->h_rhorefh.data();
    // Pres_2::set_values (src/pres_2.cxx:124-153) with the reference's own precision mix: the literals `2.`, `1.` are
    // double, so for TF = float the cosine and the products are evaluated in double from a FLOAT pi and narrowed on the
    // store; 1./(dx*dx) divides in double a product formed in TF.  (Pinned against the compiled reference:
    // tests/test_oracle_vs_ref.py::test_pres_2_glue_bitexact.)
    const TF* dz = c->h_dz.data();          // context-owned copies: the caller's metric arrays need not outlive mhh_ctx_create
    const TF* dzhi = c->h_dzhi.data();
    const TF dxidxi = (TF)(1. / (double)(TF)(g.dx * g.dx)), dyidyi = (TF)(1. / (double)(TF)(g.dy * g.dy));
    const TF pi = (TF)std::acos(-1.);
    std::vector<TF> bmati(c->nm), bmatj(g.jtot), a(g.kmax), cc(g.kmax), dz2rho(g.kmax), dz2(g.kmax);
    for (int j = 0; j < g.jtot / 2 + 1; ++j)
        bmatj[j] = (TF)(2. * (std::cos(2. * (double)pi * (double)(TF)j / (double)(TF)g.jtot) - 1.) * (double)dyidyi);
    for (int j = g.jtot / 2 + 1; j < g.jtot; ++j)
        bmatj[j] = bmatj[g.jtot - j];
    for (int i = 0; i < g.itot / 2 + 1; ++i)
        bmati[i] = (TF)(2. * (std::cos(2. * (double)pi * (double)(TF)i / (double)(TF)g.itot) - 1.) * (double)dxidxi);
    for (int k = 0; k < g.kmax; ++k)
    {
        a[k] = dz[k + g.kgc] * rh[k + g.kgc] * dzhi[k + g.kgc];
        cc[k] = dz[k + g.kgc] * rh[k + g.kgc + 1] * dzhi[k + g.kgc + 1];
        dz2[k] = dz[k + g.kgc] * dz[k + g.kgc];
        dz2rho[k] = dz2[k] * rr[k + g.kgc];
    }
    CUDA_TRY(c, cudaMemcpy(c->d_bmati, bmati.data(), sizeof(TF) * c->nm, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_bmatj, bmatj.data(), sizeof(TF) * g.jtot, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_a, a.data(), sizeof(TF) * g.kmax, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_c, cc.data(), sizeof(TF) * g.kmax, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_dz2rho, dz2rho.data(), sizeof(TF) * g.kmax, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_dz2, dz2.data(), sizeof(TF) * g.kmax, cudaMemcpyHostToDevice));

    if (c->fused)
    {
        p2_setup_launch<TF>(g.jtot, c->stream, c->fac, c->coef(), c->lay2.mcl, g.kmax, c->lay2.m_off);
        KCHECKN(c, "tdma2_setup_kernel");
    }
    else
    {
        const long long ncol = (long long)c->lay.mcl * g.jtot;
        tdma_setup_kernel<TF><<<(unsigned)((ncol + 127) / 128), 128, 0, c->stream>>>(c->fac, c->coef(), c->lay.mcl, g.jtot, g.kmax, c->lay.m_off, 0);
        KCHECKN(c, "tdma_setup_kernel");
    }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return MHH_OK;
}

// ---- Pres_2 ---------------------------------------------------------------------------------
template <typename TF>
int slab_all_to_all(Ctx<TF>* c, bool forward)
{
    if (c->nranks == 1) return MHH_OK;
    if (!c->comm) { c->err = "slab context without communicator: call mhh_comm_init first"; return MHH_E_INVALID; }
    NcclApi* api = nccl_api(c->err);
    if (!api) return MHH_E_CUDA;
    const SpecLayout& l = c->lay;
    cplx<TF>* X = reinterpret_cast<cplx<TF>*>(c->spec);
    cplx<TF>* Y = reinterpret_cast<cplx<TF>*>(c->specT);
    const size_t ybytes = sizeof(cplx<TF>) * (size_t)l.mcl * l.rows;      // every message on the y side has this size
    NCCL_TRY(c, api, api->GroupStart());
    for (int step = 1; step < l.P; ++step)
    {
        const int to = (l.rank + step) % l.P, from = (l.rank + l.P - step) % l.P;
        cplx<TF>* xb_to = X + (size_t)l.offset(to) * l.rows;
        cplx<TF>* xb_from = X + (size_t)l.offset(from) * l.rows;
        const size_t xbytes_to = sizeof(cplx<TF>) * (size_t)l.count(to) * l.rows;
        const size_t xbytes_from = sizeof(cplx<TF>) * (size_t)l.count(from) * l.rows;
        if (forward)
        {
            NCCL_TRY(c, api, api->Send(xb_to, xbytes_to, ncclChar, to, c->comm, c->stream));
            NCCL_TRY(c, api, api->Recv(Y + (size_t)from * l.mcl * l.rows, ybytes, ncclChar, from, c->comm, c->stream));
        }
        else
        {
            NCCL_TRY(c, api, api->Send(Y + (size_t)to * l.mcl * l.rows, ybytes, ncclChar, to, c->comm, c->stream));
            NCCL_TRY(c, api, api->Recv(xb_from, xbytes_from, ncclChar, from, c->comm, c->stream));
        }
    }
    NCCL_TRY(c, api, api->GroupEnd());
    cplx<TF>* xs = X + (size_t)l.offset(l.rank) * l.rows;
    cplx<TF>* ys = Y + (size_t)l.rank * l.mcl * l.rows;
    CUDA_TRY(c, cudaMemcpyAsync(forward ? ys : xs, forward ? xs : ys, ybytes, cudaMemcpyDeviceToDevice, c->stream));
    prof_mark(c, forward ? "all_to_all_xy_nccl" : "all_to_all_yx_nccl");
    return MHH_OK;
}

// solver: 0 = transforms only, 2 = tridiagonal (Pres_2), 4 = 7-band (Pres_4)
template <typename TF>
int pres_spectral_solve(Ctx<TF>* c, int solver)
{
    const GridDev<TF>& g = c->g;
    const int mcl = c->lay.mcl;
    int rc;
    // fused transposes: the x transform already stored into the owners' y-side buffers; one all-reduce is the barrier
    if ((rc = c->peers.on ? slab_barrier<TF>(c, "transpose_xy_barrier") : slab_all_to_all<TF>(c, true)) != MHH_OK) return rc;
    const int grid_p = c->num_sms * 2;
    const long long ypanels = (long long)((mcl + WFFT_WARPS - 1) / WFFT_WARPS) * g.ktot;
    const int grid_wy = (int)std::max<long long>(1, std::min<long long>(ypanels, (long long)c->num_sms * 8));
    if (g.jtot > 1)
    {
        if (c->wfft_y) wfft_y_launch<TF>(g.jtot, grid_wy, c->stream, c->specT, c->lay, c->peers, mcl, g.ktot, c->tw_y, 0);
        else fft_y_kernel<TF><<<grid_p, 256, c->smem_y, c->stream>>>(c->specT, c->lay, c->peers, mcl, g.jtot, g.ktot, c->plan_y, c->tw_y, c->mc_y, 0);
        KCHECKN(c, "fft_y_forward_kernel");
    }
    if (solver == 2)
    {
        const long long ncol = (long long)mcl * g.jtot;
        tdma_solve_kernel<TF><<<(unsigned)((ncol + 127) / 128), 128, 0, c->stream>>>(c->specT, c->fac, c->coef(), c->lay, mcl, g.jtot, g.kmax, c->lay.m_off, 0);
        KCHECKN(c, "tdma_solve_kernel");
    }
    else if (solver == 4)
    {
        const long long ncol = (long long)mcl * g.jtot;
        hdma_solve_kernel<TF><<<(unsigned)((ncol + 127) / 128), 128, 0, c->stream>>>(c->specT, c->lu4, c->lay, mcl, g.jtot, g.kmax);
        KCHECKN(c, "hdma_solve_kernel");
    }
    if (g.jtot > 1)
    {
        if (c->wfft_y) wfft_y_launch<TF>(g.jtot, grid_wy, c->stream, c->specT, c->lay, c->peers, mcl, g.ktot, c->tw_y, 1);
        else fft_y_kernel<TF><<<grid_p, 256, c->smem_y, c->stream>>>(c->specT, c->lay, c->peers, mcl, g.jtot, g.ktot, c->plan_y, c->tw_y, c->mc_y, 1);
        KCHECKN(c, "fft_y_backward_kernel");
    }
    else if (c->peers.on) { c->err = "fused transposes need jtot > 1"; return MHH_E_INVALID; }
    return c->peers.on ? slab_barrier<TF>(c, "transpose_yx_barrier") : slab_all_to_all<TF>(c, false);
}

// Pres_2 version 2: x forward -> [transpose] -> y forward + elimination -> back substitution + y inverse -> [transpose] ->
// x backward.  `compact` != NULL: the right-hand side comes from a compact (k, j, i) array (test entry point).
template <typename TF>
int pres_fused_solve(Ctx<TF>* c, const TF* compact, const RhsSrc<TF>& src, TF* p, int fill, bool do_solve)
{
    const GridDev<TF>& g = c->g;
    const Spec2& l = c->lay2;
    const int L = g.itot / 2, J = g.jtot, P = l.P;
    cplx<TF>* X = reinterpret_cast<cplx<TF>*>(c->spec);
    cplx<TF>* Y = reinterpret_cast<cplx<TF>*>(c->specT);
    NcclApi* api = nullptr;
    const bool nccl = P > 1 && !c->peers.on;
    if (P > 1 && !c->comm) { c->err = "slab context without communicator: call mhh_comm_init first"; return MHH_E_INVALID; }
    if (P > 1 && !(api = nccl_api(c->err))) return MHH_E_CUDA;
    const int npanels = l.npan * g.ktot;
    const int grid_x = std::max(1, std::min(npanels, c->num_sms * 6));
    int rc;

    // ---- x forward; its store phase is the forward transpose
    XferPtrs<TF> xf{};
    if (P == 1) xf.dst[0] = Y;
    else if (!nccl) for (int d = 0; d < P; ++d) xf.dst[d] = reinterpret_cast<cplx<TF>*>(c->peers.y[d]);
    else { xf.staged = 1; for (int d = 0; d < P; ++d) xf.dst[d] = X + (size_t)l.offset(d) * l.ktot * l.jmax; }     // the X side doubles as send staging
    p2_x_forward_launch<TF>(L, compact == nullptr, grid_x, c->stream, compact, src, g, l, xf, c->tw_xh, c->tw_xf);
    KCHECKN(c, "fft_x_forward_kernel");
    if (P > 1 && !nccl) { if ((rc = slab_barrier<TF>(c, "transpose_xy_barrier")) != MHH_OK) return rc; }
    else if (nccl)
    {
        const size_t yblk = (size_t)l.mcl * l.ktot * l.jmax;
        NCCL_TRY(c, api, api->GroupStart());
        for (int step = 1; step < P; ++step)
        {
            const int to = (l.rank + step) % P, from = (l.rank + P - step) % P;
            NCCL_TRY(c, api, api->Send(xf.dst[to], sizeof(cplx<TF>) * (size_t)l.count(to) * l.ktot * l.jmax, ncclChar, to, c->comm, c->stream));
            NCCL_TRY(c, api, api->Recv(Y + (size_t)from * yblk, sizeof(cplx<TF>) * yblk, ncclChar, from, c->comm, c->stream));
        }
        NCCL_TRY(c, api, api->GroupEnd());
        CUDA_TRY(c, cudaMemcpyAsync(Y + (size_t)l.rank * yblk, xf.dst[l.rank], sizeof(cplx<TF>) * yblk, cudaMemcpyDeviceToDevice, c->stream));
        prof_mark(c, "all_to_all_xy_nccl");
    }

    // ---- y forward + forward elimination (in place), back substitution + y inverse; its store phase is the backward transpose
    p2_y_launch<TF>(J, false, c->stream, Y, c->fac, l, xf, c->coef(), c->tw_y, c->jlog2, do_solve ? 1 : 0);
    KCHECKN(c, "fft_y_tdma_forward_kernel");
    XferPtrs<TF> xb{};
    const size_t xblk = (size_t)8 * l.mcl * l.ktot * l.npan;          // what this rank's modes contribute to ONE row owner
    if (P == 1) xb.dst[0] = X;
    else if (!nccl) for (int s = 0; s < P; ++s) xb.dst[s] = reinterpret_cast<cplx<TF>*>(c->peers.x[s]);
    else
    {
        if (!c->stage2)
        {
            CUDA_TRY(c, cudaMalloc(&c->stage2, sizeof(cplx<TF>) * xblk * P));
            c->ws_bytes += (long long)(sizeof(cplx<TF>) * xblk * P);
        }
        xb.staged = 1;
        for (int s = 0; s < P; ++s) xb.dst[s] = c->stage2 + (size_t)s * xblk;
    }
    p2_y_launch<TF>(J, true, c->stream, Y, c->fac, l, xb, c->coef(), c->tw_y, c->jlog2, do_solve ? 1 : 0);
    KCHECKN(c, "tdma_fft_y_backward_kernel");
    if (P > 1 && !nccl) { if ((rc = slab_barrier<TF>(c, "transpose_yx_barrier")) != MHH_OK) return rc; }
    else if (nccl)
    {
        NCCL_TRY(c, api, api->GroupStart());
        for (int step = 1; step < P; ++step)
        {
            const int to = (l.rank + step) % P, from = (l.rank + P - step) % P;
            NCCL_TRY(c, api, api->Send(xb.dst[to], sizeof(cplx<TF>) * xblk, ncclChar, to, c->comm, c->stream));
            NCCL_TRY(c, api, api->Recv(X + (size_t)8 * l.offset(from) * l.ktot * l.npan, sizeof(cplx<TF>) * (size_t)8 * l.count(from) * l.ktot * l.npan,
                                       ncclChar, from, c->comm, c->stream));
        }
        NCCL_TRY(c, api, api->GroupEnd());
        CUDA_TRY(c, cudaMemcpyAsync(X + (size_t)8 * l.offset(l.rank) * l.ktot * l.npan, xb.dst[l.rank], sizeof(cplx<TF>) * xblk, cudaMemcpyDeviceToDevice, c->stream));
        prof_mark(c, "all_to_all_yx_nccl");
    }

    // ---- x backward + unpack
    const TF norm = TF(1.) / ((TF)g.itot * (TF)g.jtot);
    p2_x_backward_launch<TF>(L, grid_x, c->stream, X, p, g, l, c->tw_xh, c->tw_xf, norm, fill);
    KCHECKN(c, "fft_x_backward_kernel");
    return MHH_OK;
}

template <typename TF>
int pres_solve_impl(Ctx<TF>* c, const mhh_fields* f, double sub_dt)
{
    NEED_BASE(c);
    const GridDev<TF>& g = c->g;
    NEED(c, f->p, "p");
    const bool slab = c->nranks > 1;
    if (slab)
    {
        // the divergence needs vt one row beyond the slab (src/pres_2.cxx:181 exchanges vt north-south)
        TF* vt = P<TF>(f->vt);
        int rc0 = exchange_ns<TF>(c, &vt, 1, 1, g.kcells);
        if (rc0 != MHH_OK) return rc0;
    }
    RhsSrc<TF> src{P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), P<TF>(f->ut), P<TF>(f->vt), P<TF>(f->wt), TF(1.) / (TF)sub_dt /* Pres_2::input takes dt as TF: TF(1.)/dt, src/pres_2.cxx:160,173 */, slab ? 0 : 1};
    if (c->fused)
    {
        int rcf = pres_fused_solve<TF>(c, nullptr, src, P<TF>(f->p), slab ? 0 : 1, true);
        if (rcf != MHH_OK) return rcf;
        if (slab) { TF* pp = P<TF>(f->p); return exchange_ns<TF>(c, &pp, 1, g.jgc, g.kcells); }
        return MHH_OK;
    }
    const long long nrows = (long long)g.jmax * g.ktot;
    const int grid_x = (int)std::min<long long>((nrows + c->rows_x - 1) / c->rows_x, (long long)c->num_sms * 4);
    const int grid_wx = (int)std::min<long long>((nrows + WFFT_WARPS - 1) / WFFT_WARPS, (long long)c->num_sms * 8);
    if (c->wfft_x) wfft_x_forward_launch<TF>(g.itot / 2, true, grid_wx, c->stream, c->spec, src, g, c->lay, c->peers, c->tw_xh, c->tw_xf, nrows);
    else fft_x_forward_kernel<TF, true><<<grid_x, 256, c->smem_x, c->stream>>>(c->spec, src, g, c->lay, c->peers, c->plan_x, c->tw_xh, c->tw_xf, c->rows_x, nrows);
    KCHECKN(c, "fft_x_forward_kernel");
    int rc = pres_spectral_solve<TF>(c, 2);
    if (rc != MHH_OK) return rc;
    const TF norm = TF(1.) / ((TF)g.itot * (TF)g.jtot);
    const int fill = slab ? 0 : 1;          // slabs get their north/south ghost rows of p from the neighbours below
    if (c->wfft_x) wfft_x_backward_launch<TF>(g.itot / 2, grid_wx, c->stream, c->spec, P<TF>(f->p), g, c->lay, c->tw_xh, c->tw_xf, nrows, norm, fill);
    else fft_x_backward_kernel<TF><<<grid_x, 256, c->smem_x, c->stream>>>(c->spec, P<TF>(f->p), g, c->lay, c->plan_x, c->tw_xh, c->tw_xf,
            c->rows_x, nrows, norm, fill);
    KCHECKN(c, "fft_x_backward_kernel");
    if (slab)
    {
        TF* pp = P<TF>(f->p);
        return exchange_ns<TF>(c, &pp, 1, g.jgc, g.kcells);
    }
    if (g.jtot == 1)
        return cyclic_impl<TF>(c, P<TF>(f->p), MHH_EDGE_NORTH_SOUTH, false);
    return MHH_OK;
}

template <typename TF>
int pres_exec_impl(Ctx<TF>* c, const mhh_fields* f, double sub_dt)
{
    int rc = check_mom<TF>(c, f, false, false);
    if (rc != MHH_OK) return rc;
    const GridDev<TF>& g = c->g;
    // the reference fills the east ghost cells of ut and the north ghost cells of vt as a side effect
    // (src/pres_2.cxx:180-181); keep that observable behaviour for the stand-alone entry point
    if ((rc = cyclic_impl<TF>(c, P<TF>(f->ut), MHH_EDGE_EAST_WEST, false)) != MHH_OK) return rc;
    if ((rc = cyclic_impl<TF>(c, P<TF>(f->vt), MHH_EDGE_NORTH_SOUTH, false)) != MHH_OK) return rc;
    if ((rc = pres_solve_impl<TF>(c, f, sub_dt)) != MHH_OK) return rc;
    PresArgs<TF> a{P<TF>(f->ut), P<TF>(f->vt), P<TF>(f->wt), P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), P<TF>(f->p)};
    pres_out_kernel<TF><<<c->grd_interior(), c->blk(), 0, c->stream>>>(a, g);
    KCHECKN(c, "pres_out_kernel");
    return MHH_OK;
}

// ---- Pres_4 ---------------------------------------------------------------------------------
// Pres_4::set_values (src/pres_4.cxx:178-252) + the LU factors of every mode; done at the first call
template <typename TF>
int pres4_prepare(Ctx<TF>* c)
{
    if (c->lu4) return MHH_OK;
    const GridDev<TF>& g = c->g;
    if (!g.dzi4) { c->err = "Pres_4 needs a 4th-order grid (dzi4 / dzhi4 in mhh_grid_desc, three ghost cells)"; return MHH_E_INVALID; }
    if (g.kmax < 4) { c->err = "Pres_4 needs ktot >= 4"; return MHH_E_INVALID; }
    const int kmax = g.kmax, ks = g.kstart;
    const TF dxidxi = (TF)(1. / (double)(g.dx * g.dx)), dyidyi = (TF)(1. / (double)(g.dy * g.dy));
    const double pi = (double)(TF)std::acos(-1.);
    auto wave = [&](int q, int n, TF fac) {
        return (TF)((2. * (1. / 576.) * std::cos(6. * pi * (double)q / (double)n) - 2. * (54. / 576.) * std::cos(4. * pi * (double)q / (double)n)
                   + 2. * (783. / 576.) * std::cos(2. * pi * (double)q / (double)n) - (1460. / 576.)) * (double)fac); };
    std::vector<TF> bi(c->nm), bj(g.jtot), m((size_t)7 * kmax);
    for (int i = 0; i < c->nm; ++i) bi[i] = wave(i, g.itot, dxidxi);
    for (int j = 0; j < g.jtot / 2 + 1; ++j) bj[j] = wave(j, g.jtot, dyidyi);
    for (int j = g.jtot / 2 + 1; j < g.jtot; ++j) bj[j] = bj[g.jtot - j];
    const std::vector<TF>& H = c->h_dzhi4; const std::vector<TF>& Z = c->h_dzi4;
    auto h = [&](int k) { return (double)H[k]; };
    const double f = 1. / 576.;
    auto M = [&](int n, int k) -> TF& { return m[(size_t)n * kmax + k]; };
    for (int k = 0; k < kmax; ++k)
    {
        const int kc = ks + k;
        const double z = (double)Z[kc];
        if (k == 0)
        {
            M(0, k) = 0.;
            M(1, k) = (TF)(f * (-27. * h(kc)) * z);
            M(2, k) = (TF)(f * (-1. * h(kc + 1) + 729. * h(kc) + 27. * h(kc + 1)) * z);
            M(3, k) = (TF)(f * (27. * h(kc + 1) - 729. * h(kc) - 729. * h(kc + 1) - 1. * h(kc + 2)) * z);
            M(4, k) = (TF)(f * (-27. * h(kc + 1) + 27. * h(kc) + 729. * h(kc + 1) + 27. * h(kc + 2)) * z);
            M(5, k) = (TF)(f * (1. * h(kc + 1) - 27. * h(kc + 1) - 27. * h(kc + 2)) * z);
            M(6, k) = (TF)(f * (1. * h(kc + 2)) * z);
        }
        else if (k < kmax - 1)
        {
            M(0, k) = (TF)(f * (1. * h(kc - 1)) * z);
            M(1, k) = (TF)(f * (-27. * h(kc - 1) - 27. * h(kc)) * z);
            M(2, k) = (TF)(f * (27. * h(kc - 1) + 729. * h(kc) + 27. * h(kc + 1)) * z);
            M(3, k) = (TF)(f * (-1. * h(kc - 1) - 729. * h(kc) - 729. * h(kc + 1) - 1. * h(kc + 2)) * z);
            M(4, k) = (TF)(f * (27. * h(kc) + 729. * h(kc + 1) + 27. * h(kc + 2)) * z);
            M(5, k) = (TF)(f * (-27. * h(kc + 1) - 27. * h(kc + 2)) * z);
            M(6, k) = (TF)(f * (1. * h(kc + 2)) * z);
        }
        else
        {
            M(0, k) = (TF)(f * (1. * h(kc - 1)) * z);
            M(1, k) = (TF)(f * (-27. * h(kc - 1) - 27. * h(kc) + 1. * h(kc)) * z);
            M(2, k) = (TF)(f * (27. * h(kc - 1) + 729. * h(kc) + 27. * h(kc + 1) - 27. * h(kc)) * z);
            M(3, k) = (TF)(f * (-1. * h(kc - 1) - 729. * h(kc) - 729. * h(kc + 1) + 27. * h(kc)) * z);
            M(4, k) = (TF)(f * (27. * h(kc) + 729. * h(kc + 1) - 1. * h(kc)) * z);
            M(5, k) = (TF)(f * (-27. * h(kc + 1)) * z);
            M(6, k) = 0.;
        }
    }
    const long long ncol = (long long)c->lay.mcl * g.jtot;          // this rank's modes (all of them on a single GPU)
    CUDA_TRY(c, cudaMalloc(&c->d_m7, sizeof(TF) * m.size()));
    CUDA_TRY(c, cudaMalloc(&c->d_bmati4, sizeof(TF) * bi.size()));
    CUDA_TRY(c, cudaMalloc(&c->d_bmatj4, sizeof(TF) * bj.size()));
    CUDA_TRY(c, cudaMemcpy(c->d_m7, m.data(), sizeof(TF) * m.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_bmati4, bi.data(), sizeof(TF) * bi.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_bmatj4, bj.data(), sizeof(TF) * bj.size(), cudaMemcpyHostToDevice));
    const size_t nlu = (size_t)7 * (kmax + 4) * ncol;
    CUDA_TRY(c, cudaMalloc(&c->lu4, sizeof(TF) * nlu));
    c->ws_bytes += (long long)(sizeof(TF) * nlu);
    HdmaCoef<TF> cf{c->d_m7, c->d_bmati4, c->d_bmatj4};
    hdma_setup_kernel<TF><<<(unsigned)((ncol + 127) / 128), 128, 0, c->stream>>>(c->lu4, cf, c->lay.mcl, g.jtot, kmax, c->lay.m_off);
    KCHECKN(c, "hdma_setup_kernel");
    return MHH_OK;
}

// Pres_4::exec (src/pres_4.cxx:76-144): input -> transforms + 7-band solve -> ghost cells -> output
template <typename TF>
int pres4_exec_impl(Ctx<TF>* c, const mhh_fields* f, double sub_dt)
{
    int rc = check_mom<TF>(c, f, false, false);
    if (rc != MHH_OK) return rc;
    NEED(c, f->p, "p");
    if ((rc = pres4_prepare<TF>(c)) != MHH_OK) return rc;
    const GridDev<TF>& g = c->g;
    const bool dim3 = g.jtot > 1;
    // cyclic ghosts of the tendencies and the mirrored wt over the walls are side effects of Pres_4::input
    if ((rc = cyclic_impl<TF>(c, P<TF>(f->ut), MHH_EDGE_EAST_WEST, false)) != MHH_OK) return rc;
    if (dim3 && (rc = cyclic_impl<TF>(c, P<TF>(f->vt), MHH_EDGE_NORTH_SOUTH, false)) != MHH_OK) return rc;
    {
        ::dim3 b2(64, 4), g2((g.imax + 63) / 64, (g.jmax + 3) / 4);
        pres4_wtbc_kernel<TF><<<g2, b2, 0, c->stream>>>(P<TF>(f->wt), g);
        KCHECKN(c, "pres4_wtbc_kernel");
    }
    const TF dti = (TF)(1. / (double)(TF)sub_dt);       // Pres_4::input takes dt as TF and forms 1./dt in double (src/pres_4.cxx:262,276)
    const bool slab = c->nranks > 1;
    // the right-hand side goes into the (free) p array as a compact (k, j, i) block, like the reference's pres_in does: the x
    // transform then reads it from there and its store phase is free to be the forward transpose
    TF* rhs = P<TF>(f->p);
    const long long pitch = g.itot;
    if (dim3) pres4_in_kernel<TF, true><<<c->grd_interior(), c->blk(), 0, c->stream>>>(P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), P<TF>(f->ut), P<TF>(f->vt), P<TF>(f->wt), rhs, pitch, dti, g);
    else pres4_in_kernel<TF, false><<<c->grd_interior(), c->blk(), 0, c->stream>>>(P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), P<TF>(f->ut), P<TF>(f->vt), P<TF>(f->wt), rhs, pitch, dti, g);
    KCHECKN(c, "pres4_in_kernel");
    const long long nrows = (long long)g.jmax * g.ktot;
    const int grid_x = (int)std::min<long long>((nrows + c->rows_x - 1) / c->rows_x, (long long)c->num_sms * 4);
    const int grid_wx = (int)std::min<long long>((nrows + WFFT_WARPS - 1) / WFFT_WARPS, (long long)c->num_sms * 8);
    RhsSrc<TF> from_p{};
    from_p.u = rhs;
    if (c->wfft_x) wfft_x_forward_launch<TF>(g.itot / 2, false, grid_wx, c->stream, c->spec, from_p, g, c->lay, c->peers, c->tw_xh, c->tw_xf, nrows);
    else fft_x_forward_kernel<TF, false><<<grid_x, 256, c->smem_x, c->stream>>>(c->spec, from_p, g, c->lay, c->peers, c->plan_x, c->tw_xh, c->tw_xf, c->rows_x, nrows);
    KCHECKN(c, "fft_x_forward_kernel");
    if (slab && !dim3) { c->err = "Pres_4 on y slabs needs jtot > 1"; return MHH_E_INVALID; }
    if ((rc = pres_spectral_solve<TF>(c, 4)) != MHH_OK) return rc;
    const TF norm = TF(1.) / ((TF)g.itot * (TF)g.jtot);
    const int fill = slab ? 0 : 1;
    if (c->wfft_x) wfft_x_backward_launch<TF>(g.itot / 2, grid_wx, c->stream, c->spec, P<TF>(f->p), g, c->lay, c->tw_xh, c->tw_xf, nrows, norm, fill);
    else fft_x_backward_kernel<TF><<<grid_x, 256, c->smem_x, c->stream>>>(c->spec, P<TF>(f->p), g, c->lay, c->plan_x, c->tw_xh, c->tw_xf, c->rows_x, nrows, norm, fill);
    KCHECKN(c, "fft_x_backward_kernel");
    if (slab) { TF* pp = P<TF>(f->p); if ((rc = exchange_ns<TF>(c, &pp, 1, g.jgc, g.kcells)) != MHH_OK) return rc; }
    if (!dim3 && (rc = cyclic_impl<TF>(c, P<TF>(f->p), MHH_EDGE_NORTH_SOUTH, false)) != MHH_OK) return rc;
    {
        ::dim3 b2(64, 4), g2((g.icells + 63) / 64, (g.jcells + 3) / 4);
        pres4_ghost_kernel<TF><<<g2, b2, 0, c->stream>>>(P<TF>(f->p), g);
        KCHECKN(c, "pres4_ghost_kernel");
    }
    if (dim3) pres4_out_kernel<TF, true><<<c->grd_interior(), c->blk(), 0, c->stream>>>(P<TF>(f->ut), P<TF>(f->vt), P<TF>(f->wt), P<TF>(f->p), g);
    else pres4_out_kernel<TF, false><<<c->grd_interior(), c->blk(), 0, c->stream>>>(P<TF>(f->ut), P<TF>(f->vt), P<TF>(f->wt), P<TF>(f->p), g);
    KCHECKN(c, "pres4_out_kernel");
    return MHH_OK;
}

template <typename TF>
int pres4_div_impl(Ctx<TF>* c, const mhh_fields* f, double* out)
{
    const GridDev<TF>& g = c->g;
    if (!g.dzi4) { c->err = "Pres_4 needs a 4th-order grid"; return MHH_E_INVALID; }
    NEED(c, f->u, "u"); NEED(c, f->v, "v"); NEED(c, f->w, "w");
    CUDA_TRY(c, cudaMemsetAsync(c->d_red, 0, sizeof(double), c->stream));
    pres4_div_kernel<TF><<<c->grd_interior(), c->blk(), 0, c->stream>>>(P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), g, c->d_red);
    KCHECKN(c, "pres4_div_kernel");
    if (c->nranks > 1)
    {
        if (!c->comm) { c->err = "slab context without communicator: call mhh_comm_init first"; return MHH_E_INVALID; }
        NcclApi* api = nccl_api(c->err);
        if (!api) return MHH_E_CUDA;
        NCCL_TRY(c, api, api->AllReduce(c->d_red, c->d_red, 1, ncclFloat64, ncclMax, c->comm, c->stream));
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->h_red, c->d_red, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *out = *c->h_red;
    return MHH_OK;
}

template <typename TF>
int fft_roundtrip_impl(Ctx<TF>* c, const TF* in, TF* out, int solve)
{
    NEED_BASE(c);
    const GridDev<TF>& g = c->g;
    NEED(c, in, "in"); NEED(c, out, "out");
    if (c->nranks > 1) { c->err = "pres_fft_roundtrip is a single-GPU test entry point"; return MHH_E_INVALID; }
    if (c->fused)
    {
        TF* tmp = nullptr;
        CUDA_TRY(c, cudaMalloc(&tmp, sizeof(TF) * (size_t)g.ncells));
        RhsSrc<TF> none{};
        int rcf = pres_fused_solve<TF>(c, in, none, tmp, 0, solve != 0);
        cudaError_t e = cudaSuccess;
        for (int k = 0; k < g.ktot && e == cudaSuccess && rcf == MHH_OK; ++k)
            e = cudaMemcpy2DAsync(out + (size_t)k * g.itot * g.jtot, sizeof(TF) * g.itot,
                                  tmp + g.istart + (long long)g.jstart * g.icells + (long long)(g.kstart + k) * g.ijcells,
                                  sizeof(TF) * g.icells, sizeof(TF) * g.itot, g.jtot, cudaMemcpyDeviceToDevice, c->stream);
        cudaStreamSynchronize(c->stream);
        cudaFree(tmp);
        if (rcf != MHH_OK) return rcf;
        if (e != cudaSuccess) { c->err = std::string("fft_roundtrip: ") + cudaGetErrorString(e); return MHH_E_CUDA; }
        return MHH_OK;
    }
    // stage the compact input in the workspace rows (pitch 2*nm)
    CUDA_TRY(c, cudaMemcpy2DAsync(c->spec, sizeof(TF) * 2 * c->nm, in, sizeof(TF) * g.itot, sizeof(TF) * g.itot,
                                  (size_t)g.jtot * g.ktot, cudaMemcpyDeviceToDevice, c->stream));
    const long long nrows = (long long)g.jtot * g.ktot;
    const int grid_x = (int)std::min<long long>((nrows + c->rows_x - 1) / c->rows_x, (long long)c->num_sms * 4);
    RhsSrc<TF> none{};
    const int grid_wx = (int)std::min<long long>((nrows + WFFT_WARPS - 1) / WFFT_WARPS, (long long)c->num_sms * 8);
    if (c->wfft_x) wfft_x_forward_launch<TF>(g.itot / 2, false, grid_wx, c->stream, c->spec, none, g, c->lay, c->peers, c->tw_xh, c->tw_xf, nrows);
    else fft_x_forward_kernel<TF, false><<<grid_x, 256, c->smem_x, c->stream>>>(c->spec, none, g, c->lay, c->peers, c->plan_x, c->tw_xh, c->tw_xf, c->rows_x, nrows);
    KCHECKN(c, "fft_x_forward_kernel");
    int rc = pres_spectral_solve<TF>(c, solve != 0 ? 2 : 0);
    if (rc != MHH_OK) return rc;
    // backward x into a temporary ghosted array is overkill here: use a private ghosted buffer
    TF* tmp = nullptr;
    CUDA_TRY(c, cudaMalloc(&tmp, sizeof(TF) * (size_t)g.ncells));
    const TF norm = TF(1.) / ((TF)g.itot * (TF)g.jtot);
    if (c->wfft_x) wfft_x_backward_launch<TF>(g.itot / 2, grid_wx, c->stream, c->spec, tmp, g, c->lay, c->tw_xh, c->tw_xf, nrows, norm, 0);
    else fft_x_backward_kernel<TF><<<grid_x, 256, c->smem_x, c->stream>>>(c->spec, tmp, g, c->lay, c->plan_x, c->tw_xh, c->tw_xf, c->rows_x, nrows, norm, 0);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(out, sizeof(TF) * g.itot,
                              tmp + g.istart + (long long)g.jstart * g.icells + (long long)g.kstart * g.ijcells,
                              sizeof(TF) * g.icells, sizeof(TF) * g.itot, g.jtot, cudaMemcpyDeviceToDevice, c->stream);
    // cudaMemcpy2D handles one k-slab (rows are contiguous within a slab only); loop the slabs
    for (int k = 1; k < g.ktot && e == cudaSuccess; ++k)
        e = cudaMemcpy2DAsync(out + (size_t)k * g.itot * g.jtot, sizeof(TF) * g.itot,
                              tmp + g.istart + (long long)g.jstart * g.icells + (long long)(g.kstart + k) * g.ijcells,
                              sizeof(TF) * g.icells, sizeof(TF) * g.itot, g.jtot, cudaMemcpyDeviceToDevice, c->stream);
    cudaStreamSynchronize(c->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) { c->err = std::string("fft_roundtrip: ") + cudaGetErrorString(e); return MHH_E_CUDA; }
    return MHH_OK;
}


#define INST(TF) \
    template int pres_create<TF>(Ctx<TF>*); template int pres_set_values<TF>(Ctx<TF>*); \
    template int pres_solve_impl<TF>(Ctx<TF>*, const mhh_fields*, double); template int pres_exec_impl<TF>(Ctx<TF>*, const mhh_fields*, double); \
    template int pres4_exec_impl<TF>(Ctx<TF>*, const mhh_fields*, double); template int pres4_div_impl<TF>(Ctx<TF>*, const mhh_fields*, double*); \
    template int fft_roundtrip_impl<TF>(Ctx<TF>*, const TF*, TF*, int);
INST(double)
INST(float)
#undef INST

} // namespace mhhhost
