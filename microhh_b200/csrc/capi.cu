// mhhb200 -- C ABI implementation (see include/mhhb200.h).  Host-side orchestration only:
// argument checks, launch configuration, the per-context tables.  No CPU compute path exists.
#include <cmath>
#include <cstdlib>
#include <cstdint>
#include <initializer_list>
#include <cstring>
#include <string>
#include <vector>
#include <new>

#include "../../include/mhhb200.h"
#include "common.cuh"
#include "stencil_kernels.cuh"
#include "poisson_kernels.cuh"
#include "fft_warp.cuh"
#include "tile_kernels.cuh"
#include "tile2_kernels.cuh"
#include "tile3_kernels.cuh"
#include "tile4_kernels.cuh"
#include "slab_kernels.cuh"
#include "order2_kernels.cuh"
#include "order4_kernels.cuh"
#include "pres4_kernels.cuh"
#include <cudaTypedefs.h>
#include <dlfcn.h>
#include <nccl.h>

using namespace mhh;

struct mhh_ctx
{
    int dtype = MHH_F64;
    int device = 0;
    std::string err;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    long long launches = 0;
    long long ws_bytes = 0;
    int num_sms = 148;
    // optional per-kernel timing: one event after every launch; a kernel's time is the gap to the
    // previous event on the (in-order) stream
    int tile_y = 8;             // MHH_TILE_Y=8|16: tile height of the z-marching kernels
    bool force_plain = false;   // MHH_FORCE_PLAIN=1: use the point-wise kernels everywhere (A/B comparisons)
    bool no_tma = false;        // MHH_NO_TMA=1: keep the cp.async tile kernels (A/B comparisons)
    int tile3_y = 0;            // MHH_TILE3_Y: rows per CTA of the warp-specialised kernel; 0 = 3 rows with the scalar group (13 warps), 4 without
    int tile2_y = 6;            // MHH_TILE2_Y: rows (= warps) per CTA of the TMA tile kernels
    bool fuse_scalar = true;    // MHH_FUSE_SCALAR=0: keep scalar 0 out of the momentum kernel (A/B comparisons)
    int mom_variant = 4;        // MHH_MOM=2|3|4: 2 = all components per thread, 3 = warp-specialised by component, 4 = 3 with 2 x 2 register blocks (default)
    int tile4_w = 3;            // MHH_TILE4_W=2|3: warps per component of mom4 (rows per CTA = 2x)
    int evisc_mb = 4;           // MHH_EVISC_MB=2|3|4 (measured 512^3 fp64: 2.72 | 2.51 | 2.11 ms): resident CTAs per SM the eddy-viscosity kernel is compiled for (register cap)
    int prefetch = 1;           // MHH_PREFETCH: L2 prefetch distance (levels) of the TMA tile kernels, 0 = off
    bool prof = false;
    std::vector<std::pair<const char*, cudaEvent_t>> prof_events;
    std::vector<cudaEvent_t> prof_pool;
    std::string prof_json;
    // y-slab decomposition (npx = 1, npy = nranks): NCCL communicator over the slab ranks
    int nranks = 1, rank = 0;
    ncclComm_t comm = nullptr;
    virtual ~mhh_ctx() {}
};

// ---- NCCL, bound at run time (dlopen) so that single-GPU users need no NCCL at all; inside a torch
// process this resolves to the libnccl.so.2 torch has already loaded.
struct NcclApi
{
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api(std::string& err)
{
    static NcclApi api;
    static bool tried = false;
    if (!tried)
    {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
        if (api.handle)
        {
#define NCCL_SYM(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym))
            NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); NCCL_SYM(CommInitRank, "ncclCommInitRank");
            NCCL_SYM(CommDestroy, "ncclCommDestroy"); NCCL_SYM(Send, "ncclSend"); NCCL_SYM(Recv, "ncclRecv");
            NCCL_SYM(AllReduce, "ncclAllReduce"); NCCL_SYM(GroupStart, "ncclGroupStart"); NCCL_SYM(GroupEnd, "ncclGroupEnd");
            NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
            if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Send || !api.Recv || !api.AllReduce ||
                !api.GroupStart || !api.GroupEnd || !api.GetErrorString) { dlclose(api.handle); api.handle = nullptr; }
        }
    }
    if (!api.handle) { err = "NCCL (libnccl.so.2) could not be loaded"; return nullptr; }
    return &api;
}

static void prof_mark(mhh_ctx* c, const char* name)
{
    if (!c->prof) return;
    cudaEvent_t e;
    if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
    else if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, c->stream);
    c->prof_events.emplace_back(name, e);
}

namespace {

#define CUDA_TRY(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_); return MHH_E_CUDA; } } while (0)

#define NCCL_TRY(ctx, api, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { \
    (ctx)->err = std::string(#call) + ": " + (api)->GetErrorString(r_); return MHH_E_CUDA; } } while (0)

#define KCHECKN(ctx, name) do { (ctx)->launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
    (ctx)->err = std::string("kernel launch ") + name + ": " + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + std::to_string(__LINE__); \
    return MHH_E_CUDA; } prof_mark(ctx, name); } while (0)

FftPlan make_plan(int n, bool& ok)
{
    FftPlan p; p.n = n; p.nstages = 0;
    int r = n;
    ok = true;
    const int cand[5] = {8, 4, 2, 3, 5};
    while (r > 1)
    {
        bool found = false;
        for (int c : cand)
            if (r % c == 0) { p.radix[p.nstages++] = c; r /= c; found = true; break; }
        if (!found || p.nstages >= 15) { ok = false; break; }
    }
    if (n == 1) { p.nstages = 0; }
    auto lg2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return ((1 << l) == v) ? l : -1; };
    int s = 1;
    for (int st = 0; st < p.nstages; ++st)
    {
        p.log2s[st] = lg2(s);
        p.log2nb[st] = lg2(n / p.radix[st]);
        s *= p.radix[st];
    }
    return p;
}

template <typename TF>
struct Ctx : mhh_ctx
{
    GridDev<TF> g{};
    mhh_grid_desc desc{};
    // device copies of the profiles
    TF *d_prof = nullptr;          // 12 profiles x kcells
    TF *d_mlen0 = nullptr;
    // Pres_2
    int nm = 0;
    FftPlan plan_x{}, plan_y{};
    cplx<TF> *tw_xh = nullptr, *tw_xf = nullptr, *tw_y = nullptr;
    TF *d_bmati = nullptr, *d_bmatj = nullptr, *d_a = nullptr, *d_c = nullptr, *d_dz2rho = nullptr, *d_dz2 = nullptr;
    TF *spec = nullptr;            // spectral workspace, x side: nm*jmax*ktot complex
    TF *specT = nullptr;           // y side: mcl*jtot*ktot complex (== spec on a single GPU)
    TF *fac = nullptr;             // tdma factors, mcl*jtot*ktot
    // Pres_4: band coefficients (7 x kmax), 4th-order modified wavenumbers, LU factors of every mode (7 x (kmax+4) x ncol)
    TF *d_m7 = nullptr, *d_bmati4 = nullptr, *d_bmatj4 = nullptr, *lu4 = nullptr;
    std::vector<TF> h_dzi4, h_dzhi4, h_z;
    SpecLayout lay{};
    PeerPtrs<TF> peers{};          // IPC-mapped workspaces of all slab ranks (peers.on: fused transposes)
    int *d_barrier = nullptr;      // dummy word for the all-reduce that closes a fused transpose
    // peer halos: two alternating sets of (from north, from south) receive buffers, exported over CUDA IPC; the
    // neighbours' sets are mapped in peer_halo_{south,north}
    TF *phalo = nullptr;           // own: [set][dir] x phalo_cap elements
    size_t phalo_cap = 0;
    TF *phalo_south = nullptr, *phalo_north = nullptr;     // base of the south / north neighbour's phalo
    unsigned phalo_count = 0;
    TF *halo = nullptr;            // 4 staging buffers (send south/north, recv north/south) of halo_cap elements
    size_t halo_cap = 0;
    bool basestate_set = false;
    double *d_red = nullptr;       // reduction scalar
    double *h_red = nullptr;       // pinned
    std::vector<TF> h_rhoref, h_rhorefh, h_dz;
    int rows_x = 1, mc_y = 4;
    bool wfft_x = false, wfft_y = false;   // warp-per-sequence FFT kernels (power-of-two lengths)
    size_t smem_x = 0, smem_y = 0;

    ~Ctx() override
    {
        cudaSetDevice(device);
        cudaFree(d_prof); cudaFree(d_mlen0); cudaFree(tw_xh); cudaFree(tw_xf); cudaFree(tw_y);
        cudaFree(d_bmati); cudaFree(d_bmatj); cudaFree(d_a); cudaFree(d_c); cudaFree(d_dz2rho); cudaFree(d_dz2);
        for (int r = 0; r < MAX_SLAB_RANKS; ++r)
            if (peers.on && r != rank) { if (peers.x[r]) cudaIpcCloseMemHandle(peers.x[r]); if (peers.y[r]) cudaIpcCloseMemHandle(peers.y[r]); }
        if (phalo_south && phalo_south != phalo) cudaIpcCloseMemHandle(phalo_south);
        if (phalo_north && phalo_north != phalo && phalo_north != phalo_south) cudaIpcCloseMemHandle(phalo_north);
        cudaFree(phalo);
        cudaFree(d_barrier);
        cudaFree(d_m7); cudaFree(d_bmati4); cudaFree(d_bmatj4); cudaFree(lu4);
        if (specT != spec) cudaFree(specT);
        cudaFree(spec); cudaFree(fac); cudaFree(d_red); cudaFree(halo);
        if (comm) { std::string e; NcclApi* api = nccl_api(e); if (api) api->CommDestroy(comm); }
        if (h_red) cudaFreeHost(h_red);
        if (own_stream) cudaStreamDestroy(own_stream);
    }

    TdmaCoef<TF> coef() const { return {d_a, d_c, d_dz2rho, d_dz2, d_bmati, d_bmatj}; }

    dim3 blk() const { return dim3(64, 4, 1); }
    dim3 grd_interior() const { return dim3((g.imax + 63) / 64, (g.jmax + 3) / 4, g.kmax); }
    dim3 grd_all() const { return dim3((g.icells + 63) / 64, (g.jcells + 3) / 4, g.kcells); }
};

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

template <typename TF>
int create_impl(const mhh_grid_desc* d, int dtype, int device, mhh_ctx** out)
{
    Ctx<TF>* c = new (std::nothrow) Ctx<TF>();
    if (!c) return MHH_E_NOMEM;
    *out = c;
    c->dtype = dtype; c->device = device; c->desc = *d;
    { const char* e = getenv("MHH_FORCE_PLAIN"); c->force_plain = e && e[0] == '1'; }
    { const char* e = getenv("MHH_NO_TMA"); c->no_tma = e && e[0] == '1'; }
    { const char* e = getenv("MHH_TILE2_Y"); if (e) { int v = atoi(e); if (v == 4 || v == 6 || v == 8 || v == 12) c->tile2_y = v; } }
    { const char* e = getenv("MHH_FUSE_SCALAR"); if (e) c->fuse_scalar = e[0] == '1'; }
    { const char* e = getenv("MHH_MOM"); if (e && (atoi(e) == 2 || atoi(e) == 3)) c->mom_variant = atoi(e); }
    { const char* e = getenv("MHH_TILE4_W"); if (e && (atoi(e) == 2 || atoi(e) == 3)) c->tile4_w = atoi(e); }
    { const char* e = getenv("MHH_TILE3_Y"); if (e) { int v = atoi(e); if (v == 3 || v == 4 || v == 5) c->tile3_y = v; } }
    { const char* e = getenv("MHH_EVISC_MB"); if (e) c->evisc_mb = atoi(e); }
    { const char* e = getenv("MHH_PREFETCH"); if (e) c->prefetch = std::max(0, std::min(8, atoi(e))); }
    { const char* e = getenv("MHH_TILE_Y"); if (e && atoi(e) == 16) c->tile_y = 16; else if (e && atoi(e) == 8) c->tile_y = 8; }
    CUDA_TRY(c, cudaSetDevice(device));
    int nsm = 0;
    CUDA_TRY(c, cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    c->num_sms = nsm;
    CUDA_TRY(c, cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;

    GridDev<TF>& g = c->g;
    g.itot = d->itot; g.jtot = d->jtot; g.ktot = d->ktot;
    g.imax = d->imax; g.jmax = d->jmax; g.kmax = d->kmax;
    g.igc = d->igc; g.jgc = d->jgc; g.kgc = d->kgc;
    g.icells = g.imax + 2 * g.igc; g.jcells = g.jmax + 2 * g.jgc; g.kcells = g.kmax + 2 * g.kgc;
    g.istart = g.igc; g.iend = g.igc + g.imax;
    g.jstart = g.jgc; g.jend = g.jgc + g.jmax;
    g.kstart = g.kgc; g.kend = g.kgc + g.kmax;
    g.ijcells = (long long)g.icells * g.jcells;
    g.ncells = g.ijcells * g.kcells;
    // src/grid.cxx:250-253: dx = xsize/itot in TF, dxi = 1/dx
    g.dx = (TF)((TF)d->xsize / (TF)d->itot);
    g.dy = (TF)((TF)d->ysize / (TF)d->jtot);
    g.dxi = TF(1.) / g.dx;
    g.dyi = TF(1.) / g.dy;
    g.zsize = (TF)d->zsize;

    // y slabs: x and z are never split (one all-to-all pair per Poisson solve instead of the pencil layout's three)
    const int P = d->npy;
    if (d->npx != 1 || P < 1)
    { c->err = "the decomposition is y slabs: npx must be 1 and npy >= 1"; return MHH_E_INVALID; }
    if (g.jtot % P != 0 || g.imax != g.itot || g.jmax != g.jtot / P || g.kmax != g.ktot)
    { c->err = "need imax = itot, jmax = jtot/npy, kmax = ktot"; return MHH_E_INVALID; }
    if (d->mpicoordx != 0 || d->mpicoordy < 0 || d->mpicoordy >= P) { c->err = "mpicoordy out of range"; return MHH_E_INVALID; }
    if (P > 1 && (g.jmax < g.jgc || g.itot / 2 + 1 < P)) { c->err = "slab too thin for this many ranks"; return MHH_E_INVALID; }
    c->nranks = P; c->rank = d->mpicoordy;
    c->lay = make_spec_layout(g.itot, g.jtot, g.ktot, P, c->rank);
    if (g.kmax < 6) { c->err = "ktot must be >= 6"; return MHH_E_INVALID; }
    if (g.igc < 1 || g.kgc < 1 || g.jgc < 1) { c->err = "need at least one ghost cell"; return MHH_E_INVALID; }
    if (g.itot % 2 != 0) { c->err = "itot must be even"; return MHH_E_INVALID; }

    const int kc = g.kcells;
    CUDA_TRY(c, cudaMalloc(&c->d_prof, sizeof(TF) * kc * 12));
    CUDA_TRY(c, cudaMemset(c->d_prof, 0, sizeof(TF) * kc * 12));
    const void* src[6] = {d->z, d->zh, d->dz, d->dzh, d->dzi, d->dzhi};
    for (int n = 0; n < 6; ++n)
    {
        if (!src[n]) { c->err = "grid metric array is NULL"; return MHH_E_INVALID; }
        CUDA_TRY(c, cudaMemcpy(c->d_prof + n * kc, src[n], sizeof(TF) * kc, cudaMemcpyHostToDevice));
    }
    c->h_dz.assign(static_cast<const TF*>(d->dz), static_cast<const TF*>(d->dz) + kc);
    c->h_z.assign(static_cast<const TF*>(d->z), static_cast<const TF*>(d->z) + kc);
    g.z = c->d_prof; g.zh = c->d_prof + kc; g.dz = c->d_prof + 2 * kc; g.dzh = c->d_prof + 3 * kc;
    g.dzi = c->d_prof + 4 * kc; g.dzhi = c->d_prof + 5 * kc;
    g.rhoref = c->d_prof + 6 * kc; g.rhorefh = c->d_prof + 7 * kc;
    g.thref = c->d_prof + 8 * kc; g.threfh = c->d_prof + 9 * kc;
    g.dzi4 = nullptr; g.dzhi4 = nullptr;
    if (d->dzi4 && d->dzhi4)
    {
        // 4th-order grid (src/grid.cxx:306-375): three ghost cells everywhere
        if (g.igc < 3 || g.jgc < 3 || g.kgc < 3) { c->err = "a 4th-order grid needs igc, jgc, kgc >= 3"; return MHH_E_INVALID; }
        CUDA_TRY(c, cudaMemcpy(c->d_prof + 10 * kc, d->dzi4, sizeof(TF) * kc, cudaMemcpyHostToDevice));
        CUDA_TRY(c, cudaMemcpy(c->d_prof + 11 * kc, d->dzhi4, sizeof(TF) * kc, cudaMemcpyHostToDevice));
        g.dzi4 = c->d_prof + 10 * kc; g.dzhi4 = c->d_prof + 11 * kc;
        c->h_dzi4.assign(static_cast<const TF*>(d->dzi4), static_cast<const TF*>(d->dzi4) + kc);
        c->h_dzhi4.assign(static_cast<const TF*>(d->dzhi4), static_cast<const TF*>(d->dzhi4) + kc);
    }

    CUDA_TRY(c, cudaMalloc(&c->d_barrier, sizeof(int)));
    CUDA_TRY(c, cudaMemset(c->d_barrier, 0, sizeof(int)));
    CUDA_TRY(c, cudaMalloc(&c->d_red, sizeof(double)));
    CUDA_TRY(c, cudaMallocHost(&c->h_red, sizeof(double)));

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

    // x side: room for the 8-mode-panel layout of the fused peer transposes (a few per cent of padding when P > 1)
    SpecLayout tiled = c->lay; tiled.xtiled = 1;
    const size_t nspec = (size_t)2 * std::max<long long>((long long)c->nm * g.jmax * g.ktot, P > 1 ? tiled.xside_elems() : 0);
    const size_t nspecT = (size_t)2 * c->lay.mcl * g.jtot * g.ktot;
    const size_t nfac = (size_t)c->lay.mcl * g.jtot * g.ktot;
    CUDA_TRY(c, cudaMalloc(&c->spec, sizeof(TF) * nspec));
    if (P > 1) CUDA_TRY(c, cudaMalloc(&c->specT, sizeof(TF) * nspecT)); else c->specT = c->spec;
    CUDA_TRY(c, cudaMalloc(&c->fac, sizeof(TF) * nfac));
    c->ws_bytes = (long long)(sizeof(TF) * (nspec + (P > 1 ? nspecT : 0) + nfac));
    CUDA_TRY(c, cudaMalloc(&c->d_bmati, sizeof(TF) * c->nm));
    CUDA_TRY(c, cudaMalloc(&c->d_bmatj, sizeof(TF) * g.jtot));
    CUDA_TRY(c, cudaMalloc(&c->d_a, sizeof(TF) * g.kmax));
    CUDA_TRY(c, cudaMalloc(&c->d_c, sizeof(TF) * g.kmax));
    CUDA_TRY(c, cudaMalloc(&c->d_dz2rho, sizeof(TF) * g.kmax));
    CUDA_TRY(c, cudaMalloc(&c->d_dz2, sizeof(TF) * g.kmax));
    CUDA_TRY(c, cudaMalloc(&c->d_mlen0, sizeof(TF) * kc));

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
    return MHH_OK;
}

template <typename TF>
int set_basestate_impl(Ctx<TF>* c, const void* rhoref, const void* rhorefh, const void* thref, const void* threfh)
{
    GridDev<TF>& g = c->g;
    const int kc = g.kcells;
    if (!rhoref || !rhorefh) { c->err = "rhoref/rhorefh must not be NULL"; return MHH_E_INVALID; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaMemcpy(const_cast<TF*>(g.rhoref), rhoref, sizeof(TF) * kc, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(const_cast<TF*>(g.rhorefh), rhorefh, sizeof(TF) * kc, cudaMemcpyHostToDevice));
    if (thref) CUDA_TRY(c, cudaMemcpy(const_cast<TF*>(g.thref), thref, sizeof(TF) * kc, cudaMemcpyHostToDevice));
    if (threfh) CUDA_TRY(c, cudaMemcpy(const_cast<TF*>(g.threfh), threfh, sizeof(TF) * kc, cudaMemcpyHostToDevice));
    const TF* rr = static_cast<const TF*>(rhoref);
    const TF* rh = static_cast<const TF*>(rhorefh);
    c->h_rhoref.assign(rr, rr + kc); c->h_rhorefh.assign(rh, rh + kc);

    // Pres_2::set_values (src/pres_2.cxx:124-153), in TF arithmetic like the reference
    const mhh_grid_desc& d = c->desc;
    const TF* dz = static_cast<const TF*>(d.dz);
    const TF* dzhi = static_cast<const TF*>(d.dzhi);
    const TF dxidxi = TF(1.) / (g.dx * g.dx), dyidyi = TF(1.) / (g.dy * g.dy);
    const TF pi = std::acos(TF(-1.));
    std::vector<TF> bmati(c->nm), bmatj(g.jtot), a(g.kmax), cc(g.kmax), dz2rho(g.kmax), dz2(g.kmax), mlen0(kc, TF(0));
    for (int j = 0; j < g.jtot / 2 + 1; ++j)
        bmatj[j] = TF(2.) * (std::cos(TF(2.) * pi * (TF)j / (TF)g.jtot) - TF(1.)) * dyidyi;
    for (int j = g.jtot / 2 + 1; j < g.jtot; ++j)
        bmatj[j] = bmatj[g.jtot - j];
    for (int i = 0; i < g.itot / 2 + 1; ++i)
        bmati[i] = TF(2.) * (std::cos(TF(2.) * pi * (TF)i / (TF)g.itot) - TF(1.)) * dxidxi;
    for (int k = 0; k < g.kmax; ++k)
    {
        a[k] = dz[k + g.kgc] * rh[k + g.kgc] * dzhi[k + g.kgc];
        cc[k] = dz[k + g.kgc] * rh[k + g.kgc + 1] * dzhi[k + g.kgc + 1];
        dz2[k] = dz[k + g.kgc] * dz[k + g.kgc];
        dz2rho[k] = dz2[k] * rr[k + g.kgc];
    }
    // Smagorinsky filter width per level: mlen0 = (dx*dy*dz)^(1/3) (cs applied at call time)
    for (int k = 0; k < kc; ++k)
        mlen0[k] = std::pow(g.dx * g.dy * dz[k], TF(1. / 3.));
    CUDA_TRY(c, cudaMemcpy(c->d_bmati, bmati.data(), sizeof(TF) * c->nm, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_bmatj, bmatj.data(), sizeof(TF) * g.jtot, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_a, a.data(), sizeof(TF) * g.kmax, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_c, cc.data(), sizeof(TF) * g.kmax, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_dz2rho, dz2rho.data(), sizeof(TF) * g.kmax, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_dz2, dz2.data(), sizeof(TF) * g.kmax, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_mlen0, mlen0.data(), sizeof(TF) * kc, cudaMemcpyHostToDevice));

    const long long ncol = (long long)c->lay.mcl * g.jtot;
    tdma_setup_kernel<TF><<<(unsigned)((ncol + 127) / 128), 128, 0, c->stream>>>(c->fac, c->coef(), c->lay.mcl, g.jtot, g.kmax, c->lay.m_off, 0);
    KCHECKN(c, "tdma_setup_kernel");
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->basestate_set = true;
    return MHH_OK;
}

template <typename TF> inline TF* P(void* p) { return static_cast<TF*>(p); }
template <typename TF> inline const TF* P(const void* p) { return static_cast<const TF*>(p); }

#define NEED_BASE(c) do { if (!(c)->basestate_set) { (c)->err = "mhh_set_basestate has not been called"; return MHH_E_INVALID; } } while (0)
#define NEED(c, ptr, what) do { if (!(ptr)) { (c)->err = std::string(what) + " is NULL"; return MHH_E_INVALID; } } while (0)

template <typename TF>
int slab_barrier(Ctx<TF>* c, const char* name);

// north/south ghost rows of a batch of fields from the slab neighbours (periodic in y across ranks)
template <typename TF>
int exchange_ns(Ctx<TF>* c, TF* const* flds, int nf, int w, int nk)
{
    const GridDev<TF>& g0 = c->g;
    if (nf < 1 || nf > HALO_MAX_FIELDS || w < 1 || w > g0.jgc) { c->err = "exchange_ns: bad batch"; return MHH_E_INVALID; }
    if (!c->comm) { c->err = "slab context without communicator: call mhh_comm_init first"; return MHH_E_INVALID; }
    NcclApi* api = nccl_api(c->err);
    if (!api) return MHH_E_CUDA;
    GridDev<TF> g = g0;
    g.kcells = nk;                                  // 2-D companions: one level
    const size_t per = (size_t)w * g.icells * nk;
    const size_t need = per * nf;
    if (c->phalo_south && need <= c->phalo_cap)
    {
        // peer halos: push the strips into the neighbours' receive buffers, barrier, unpack the own ones.  Two buffer sets
        // alternate so that a neighbour that is still unpacking exchange n is never overwritten by exchange n+1.
        HaloFields<TF> h{}; h.nf = nf;
        for (int n = 0; n < nf; ++n) { NEED(c, flds[n], "field"); h.f[n] = flds[n]; }
        const size_t set = (size_t)(c->phalo_count++ & 1u) * 2 * c->phalo_cap;
        const int grid = (int)std::min<size_t>((need + 255) / 256, (size_t)c->num_sms * 8);
        halo_push_kernel<TF><<<grid, 256, 0, c->stream>>>(h, g, w, c->phalo_south + set, c->phalo_north + set + c->phalo_cap);
        KCHECKN(c, "halo_push_kernel");
        int rcb = slab_barrier<TF>(c, "halo_barrier");
        if (rcb != MHH_OK) return rcb;
        halo_unpack_kernel<TF><<<grid, 256, 0, c->stream>>>(h, g, w, c->phalo + set, c->phalo + set + c->phalo_cap);
        KCHECKN(c, "halo_unpack_kernel");
        return MHH_OK;
    }
    if (c->halo_cap < need)
    {
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->halo); c->halo = nullptr; c->halo_cap = 0;
        CUDA_TRY(c, cudaMalloc(&c->halo, sizeof(TF) * need * 4));
        c->halo_cap = need;
    }
    TF* sendS = c->halo; TF* sendN = c->halo + c->halo_cap; TF* recvN = c->halo + 2 * c->halo_cap; TF* recvS = c->halo + 3 * c->halo_cap;
    HaloFields<TF> h{}; h.nf = nf;
    for (int n = 0; n < nf; ++n) { NEED(c, flds[n], "field"); h.f[n] = flds[n]; }
    const int grid = (int)std::min<size_t>((need + 255) / 256, (size_t)c->num_sms * 8);
    halo_pack_kernel<TF><<<grid, 256, 0, c->stream>>>(h, g, w, sendS, sendN);
    KCHECKN(c, "halo_pack_kernel");
    const int south = (c->rank + c->nranks - 1) % c->nranks, north = (c->rank + 1) % c->nranks;
    const size_t bytes = need * sizeof(TF);
    NCCL_TRY(c, api, api->GroupStart());
    NCCL_TRY(c, api, api->Send(sendS, bytes, ncclChar, south, c->comm, c->stream));
    NCCL_TRY(c, api, api->Send(sendN, bytes, ncclChar, north, c->comm, c->stream));
    NCCL_TRY(c, api, api->Recv(recvN, bytes, ncclChar, north, c->comm, c->stream));
    NCCL_TRY(c, api, api->Recv(recvS, bytes, ncclChar, south, c->comm, c->stream));
    NCCL_TRY(c, api, api->GroupEnd());
    prof_mark(c, "halo_sendrecv_nccl");
    halo_unpack_kernel<TF><<<grid, 256, 0, c->stream>>>(h, g, w, recvN, recvS);
    KCHECKN(c, "halo_unpack_kernel");
    return MHH_OK;
}

template <typename TF>
int cyclic_local(Ctx<TF>* c, TF* fld, int edge, bool two_d);

// Boundary_cyclic::exec on one field: local periodic copies, plus the neighbour exchange for y slabs
template <typename TF>
int cyclic_impl(Ctx<TF>* c, TF* fld, int edge, bool two_d)
{
    if (c->nranks == 1) return cyclic_local<TF>(c, fld, edge, two_d);
    if (edge < 0 || edge > 2) { c->err = "bad edge"; return MHH_E_INVALID; }
    int rc;
    if (edge != MHH_EDGE_NORTH_SOUTH && (rc = cyclic_local<TF>(c, fld, MHH_EDGE_EAST_WEST, two_d)) != MHH_OK) return rc;
    if (edge == MHH_EDGE_EAST_WEST) return MHH_OK;
    return exchange_ns<TF>(c, &fld, 1, c->g.jgc, two_d ? 1 : c->g.kcells);
}

// a batch of fields: one message per direction for all of them
template <typename TF>
int cyclic_fields(Ctx<TF>* c, TF* const* flds, int nf)
{
    int rc;
    for (int n = 0; n < nf; ++n)
        if ((rc = cyclic_local<TF>(c, flds[n], c->nranks == 1 ? MHH_EDGE_BOTH : MHH_EDGE_EAST_WEST, false)) != MHH_OK) return rc;
    if (c->nranks == 1) return MHH_OK;
    return exchange_ns<TF>(c, flds, nf, c->g.jgc, c->g.kcells);
}

template <typename TF>
int cyclic_local(Ctx<TF>* c, TF* fld, int edge, bool two_d)
{
    const GridDev<TF>& g = c->g;
    NEED(c, fld, "field");
    if (edge < 0 || edge > 2) { c->err = "bad edge"; return MHH_E_INVALID; }
    const int nk = two_d ? 1 : g.kcells;
    const int n0 = 2 * g.igc * g.jcells, n1 = 2 * g.jgc * g.icells;
    const int nmax = std::max(n0, n1);
    dim3 grid((nmax + 255) / 256, nk, 2);
    cyclic_kernel<TF><<<grid, 256, 0, c->stream>>>(fld, g, edge, nk);
    KCHECKN(c, "cyclic_kernel");
    return MHH_OK;
}

template <typename TF>
int ghost_impl(Ctx<TF>* c, TF* fld, int bcbot, const TF* bot, const TF* gradbot, int bctop, const TF* top, const TF* gradtop)
{
    const GridDev<TF>& g = c->g;
    NEED(c, fld, "field");
    if (bcbot == MHH_BC_DIRICHLET) NEED(c, bot, "bot");
    if (bcbot == MHH_BC_NEUMANN) NEED(c, gradbot, "gradbot");
    if (bctop == MHH_BC_DIRICHLET) NEED(c, top, "top");
    if (bctop == MHH_BC_NEUMANN) NEED(c, gradtop, "gradtop");
    dim3 b(64, 4), gr((g.icells + 63) / 64, (g.jcells + 3) / 4);
    ghost_cells_2nd_kernel<TF><<<gr, b, 0, c->stream>>>(fld, g, bcbot, bot, gradbot, bctop, top, gradtop);
    KCHECKN(c, "ghost_cells_2nd_kernel");
    return MHH_OK;
}

template <typename TF>
int vec_width(const GridDev<TF>& g, std::initializer_list<const void*> ptrs)
{
    int v = 2;
    if (g.icells % 2 != 0 || (g.igc - TILE_H) % 2 != 0 || (g.ijcells % 2) != 0) v = 1;
    for (const void* p : ptrs)
        if (p && (reinterpret_cast<uintptr_t>(p) % (2 * sizeof(TF))) != 0) v = 1;
    return v;
}

inline int pick_kchunk_waves(int ntiles_xy, int kmax, int slots, int warm);

inline int pick_kchunk(int ntiles_xy, int kmax, int num_sms)
{
    // enough CTAs to fill the machine twice, but chunks of at least 16 levels (warm-up level amortised)
    int nz = (2 * num_sms + ntiles_xy - 1) / ntiles_xy;
    nz = std::max(1, std::min(nz, std::max(1, kmax / 16)));
    return (kmax + nz - 1) / nz;
}

template <typename TF>
int evisc_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, const TF* n2)
{
    NEED_BASE(c);
    const GridDev<TF>& g = c->g;
    NEED(c, f->evisc, "evisc"); NEED(c, f->u, "u"); NEED(c, f->v, "v"); NEED(c, f->w, "w");
    EviscArgs<TF> a{};
    a.evisc = P<TF>(f->evisc); a.u = P<TF>(f->u); a.v = P<TF>(f->v); a.w = P<TF>(f->w);
    a.n2 = n2; a.th = nullptr;
    // Thermo_type::Disabled (src/diff_smag2.cxx:507-545): no stability correction, calc_evisc_neutral
    const bool neutral = !n2 && prm->swthermo == 0;
    if (!n2 && !neutral)
    {
        if (f->ns < 1 || !f->s[0]) { c->err = "exec_viscosity: no N2 field and no scalar 0 (th) to derive it from"; return MHH_E_INVALID; }
        a.th = P<TF>(f->s[0]);
    }
    a.n2mode = n2 ? 0 : 1;
    a.surface = prm->surface_model; a.mason = prm->sw_mason;
    a.cs = (TF)prm->cs; a.tPr = (TF)prm->tPr;
    if (a.surface)
    {
        NEED(c, f->dudz_mo, "dudz_mo"); NEED(c, f->dvdz_mo, "dvdz_mo"); NEED(c, f->z0m, "z0m");
        if (!neutral) NEED(c, f->dbdz_mo, "dbdz_mo");
        a.dudz = P<TF>(f->dudz_mo); a.dvdz = P<TF>(f->dvdz_mo); a.dbdz = P<TF>(f->dbdz_mo); a.z0m = P<TF>(f->z0m);
    }
    if (neutral)
    {
        evisc_neutral_kernel<TF><<<c->grd_interior(), c->blk(), 0, c->stream>>>(a, g, c->d_mlen0, (TF)f->visc);
        KCHECKN(c, "evisc_neutral_kernel");
    }
    else if (!c->force_plain)
    {
        const int ty = c->tile_y;
        const int ntx = (g.imax + TILE_X - 1) / TILE_X, nty = (g.jmax + ty - 1) / ty;
        const int mb = (ty == 16) ? 1 : (c->evisc_mb == 3 || c->evisc_mb == 4 ? c->evisc_mb : 2);
        EviscTileArgs<TF> t{a, c->d_mlen0, pick_kchunk_waves(ntx * nty, g.kmax, c->num_sms * mb, 1)};
        dim3 grid(ntx, nty, (g.kmax + t.kchunk - 1) / t.kchunk);
        const size_t smem = evisc_tile_smem(sizeof(TF), t.kchunk, ty);
        int vec = 2;
        if (g.icells % 2 != 0 || (g.igc - EH) % 2 != 0 || (g.ijcells % 2) != 0) vec = 1;
        for (const void* p : {(const void*)a.u, (const void*)a.v, (const void*)a.w})
            if (reinterpret_cast<uintptr_t>(p) % (2 * sizeof(TF)) != 0) vec = 1;
#define ET4(S, V, Y, MB) do { \
            static size_t attr_smem_dev[64] = {0}; size_t& attr_smem = attr_smem_dev[c->device & 63];   /* the attribute is per device */ \
            if (attr_smem < smem) { CUDA_TRY(c, cudaFuncSetAttribute(evisc_tile_kernel<TF, S, V, Y, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; } \
            evisc_tile_kernel<TF, S, V, Y, MB><<<grid, TILE_X * Y, smem, c->stream>>>(t, g); } while (0)
#define ET3(S, V, Y) do { if (c->evisc_mb == 3) ET4(S, V, Y, 3); else if (c->evisc_mb == 4) ET4(S, V, Y, 4); else ET4(S, V, Y, 512 / (TILE_X * Y)); } while (0)
#define ET(S, V) do { if (ty == 16) ET4(S, V, 16, 1); else ET3(S, V, 8); } while (0)
        if (a.surface) { if (vec == 2) ET(true, 2); else ET(true, 1); }
        else { if (vec == 2) ET(false, 2); else ET(false, 1); }
#undef ET
#undef ET3
#undef ET4
        KCHECKN(c, "evisc_tile_kernel");
    }
    else
    {
        evisc_kernel<TF><<<c->grd_interior(), c->blk(), 0, c->stream>>>(a, g, c->d_mlen0);
        KCHECKN(c, "evisc_kernel");
    }
    if (!a.surface)
    {
        dim3 b(64, 4), gr((g.icells + 63) / 64, (g.jcells + 3) / 4);
        evisc_mirror_kernel<TF><<<gr, b, 0, c->stream>>>(a.evisc, g);
        KCHECKN(c, "evisc_mirror_kernel");
    }
    return cyclic_impl<TF>(c, a.evisc, MHH_EDGE_BOTH, false);
}
template <typename TF>
MomArgs<TF> mom_args(const mhh_fields* f)
{
    MomArgs<TF> a{};
    a.ut = P<TF>(f->ut); a.vt = P<TF>(f->vt); a.wt = P<TF>(f->wt);
    a.u = P<TF>(f->u); a.v = P<TF>(f->v); a.w = P<TF>(f->w);
    a.evisc = P<TF>(f->evisc);
    a.th = (f->ns > 0) ? P<TF>(f->s[0]) : nullptr;
    a.u_fluxbot = P<TF>(f->u_fluxbot); a.u_fluxtop = P<TF>(f->u_fluxtop);
    a.v_fluxbot = P<TF>(f->v_fluxbot); a.v_fluxtop = P<TF>(f->v_fluxtop);
    a.visc = (TF)f->visc;
    return a;
}

template <typename TF>
ScalArgs<TF> scal_args(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, int n)
{
    const GridDev<TF>& g = c->g;
    ScalArgs<TF> a{};
    a.st = P<TF>(f->st[n]); a.s = P<TF>(f->s[n]);
    a.u = P<TF>(f->u); a.v = P<TF>(f->v); a.w = P<TF>(f->w);
    a.evisc = P<TF>(f->evisc);
    a.fluxbot = P<TF>(f->s_fluxbot[n]); a.fluxtop = P<TF>(f->s_fluxtop[n]);
    a.visc = (TF)f->svisc[n];
    a.tPr = prm ? (TF)prm->tPr : TF(1);
    // 1./(dx*dx) is formed in double in the reference and narrowed to TF (src/diff_smag2.cxx:445)
    a.dxidxi = (TF)(1. / ((double)g.dx * (double)g.dx));
    a.dyidyi = (TF)(1. / ((double)g.dy * (double)g.dy));
    return a;
}

template <typename TF>
int check_mom(Ctx<TF>* c, const mhh_fields* f, bool need_evisc, bool surface)
{
    NEED(c, f, "fields");
    NEED(c, f->u, "u"); NEED(c, f->v, "v"); NEED(c, f->w, "w");
    NEED(c, f->ut, "ut"); NEED(c, f->vt, "vt"); NEED(c, f->wt, "wt");
    if (f->ns < 0 || f->ns > MHH_MAX_SCALARS) { c->err = "ns out of range"; return MHH_E_INVALID; }
    for (int n = 0; n < f->ns; ++n) { NEED(c, f->s[n], "scalar"); NEED(c, f->st[n], "scalar tendency"); }
    if (need_evisc) NEED(c, f->evisc, "evisc");
    if (need_evisc && surface)
    {
        NEED(c, f->u_fluxbot, "u_fluxbot"); NEED(c, f->u_fluxtop, "u_fluxtop");
        NEED(c, f->v_fluxbot, "v_fluxbot"); NEED(c, f->v_fluxtop, "v_fluxtop");
        for (int n = 0; n < f->ns; ++n) { NEED(c, f->s_fluxbot[n], "s_fluxbot"); NEED(c, f->s_fluxtop[n], "s_fluxtop"); }
    }
    return MHH_OK;
}

// ---- TMA tensor maps (driver entry point fetched through the runtime; libcuda is not linked) ----
inline PFN_cuTensorMapEncodeTiled tmap_encoder()
{
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    if (!fn)
    {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    }
    return fn;
}

// 3-D map over a ghosted field (icells, jcells, kcells) with a box of (bx, by, 1) elements
template <typename TF>
bool make_field_tmap(CUtensorMap* m, const void* fld, const GridDev<TF>& g, int bx, int by)
{
    PFN_cuTensorMapEncodeTiled enc = tmap_encoder();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)g.icells, (cuuint64_t)g.jcells, (cuuint64_t)g.kcells};
    const cuuint64_t strides[2] = {(cuuint64_t)g.icells * sizeof(TF), (cuuint64_t)g.ijcells * sizeof(TF)};
    const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUtensorMapDataType dt = sizeof(TF) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    return enc(m, dt, 3, const_cast<void*>(fld), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// TMA needs 16-byte aligned base pointers and row/plane pitches that are multiples of 16 bytes
template <typename TF>
bool tma_ok(const GridDev<TF>& g, std::initializer_list<const void*> ptrs)
{
    if (((size_t)g.icells * sizeof(TF)) % 16 != 0 || (g.imax % 2) != 0 || g.igc < 3 || g.jgc < 3) return false;
    for (const void* p : ptrs)
        if (!p || (reinterpret_cast<uintptr_t>(p) % 16) != 0) return false;
    return tmap_encoder() != nullptr;
}

// z-chunk of the marching kernels: fill whole waves of resident CTAs, pay (warm-up levels)/kchunk per chunk
inline int pick_kchunk_waves(int ntiles_xy, int kmax, int slots, int warm)
{
    int best_nz = 1; double best = -1.;
    for (int nz = 1; nz <= std::max(1, kmax / 16); ++nz)
    {
        const int kchunk = (kmax + nz - 1) / nz;
        const int nzz = (kmax + kchunk - 1) / kchunk;
        const long long ctas = (long long)ntiles_xy * nzz;
        const long long waves = (ctas + slots - 1) / slots;
        const double eff = (double)ctas / (double)(waves * slots) * (double)kmax / (double)(nzz * (kchunk + warm));
        if (eff > best * 1.02) { best = eff; best_nz = nz; }
    }
    return (kmax + best_nz - 1) / best_nz;
}

// fused advection + diffusion (+ buoyancy) of u, v, w with the z-marching tile kernel
template <typename TF>
int mom_tile_launch(Ctx<TF>* c, const MomArgs<TF>& a, bool surface, bool buoy)
{
    const GridDev<TF>& g = c->g;
    const int ty = c->tile_y;
    const int ntx = (g.imax + TILE_X - 1) / TILE_X, nty = (g.jmax + ty - 1) / ty;
    MomTileArgs<TF> t{a, pick_kchunk(ntx * nty, g.kmax, c->num_sms)};
    dim3 grid(ntx, nty, (g.kmax + t.kchunk - 1) / t.kchunk);
    const size_t smem = mom_tile_smem(sizeof(TF), t.kchunk, ty);
    const int vec = vec_width<TF>(g, {a.u, a.v, a.w, a.evisc});
#define MT(S, B, V, Y) do { \
        static size_t attr_smem_dev[64] = {0}; size_t& attr_smem = attr_smem_dev[c->device & 63];   /* the attribute is per device */ \
        if (attr_smem < smem) { CUDA_TRY(c, cudaFuncSetAttribute(mom_tile_kernel<TF, S, B, V, Y>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; } \
        mom_tile_kernel<TF, S, B, V, Y><<<grid, TILE_X * Y, smem, c->stream>>>(t, g); } while (0)
#define MT2(S, B) do { if (vec == 2) { if (ty == 16) MT(S, B, 2, 16); else MT(S, B, 2, 8); } \
                       else { if (ty == 16) MT(S, B, 1, 16); else MT(S, B, 1, 8); } } while (0)
    if (surface && buoy) MT2(true, true);
    else if (surface) MT2(true, false);
    else if (buoy) MT2(false, true);
    else MT2(false, false);
#undef MT2
#undef MT
    KCHECKN(c, "mom_tile_kernel");
    return MHH_OK;
}

// TMA-staged, two-columns-per-thread variant (tile2_kernels.cuh)
template <typename TF>
int mom2_launch(Ctx<TF>* c, const MomArgs<TF>& a, bool surface, bool buoy)
{
    const GridDev<TF>& g = c->g;
    const int ty = c->tile2_y;
    const int ntx = (g.imax + T2_W - 1) / T2_W, nty = (g.jmax + ty - 1) / ty;
    const int per_sm = ty <= 8 ? 2 : 1;
    Mom2Args<TF> t{a, pick_kchunk_waves(ntx * nty, g.kmax, c->num_sms * per_sm, 2), c->prefetch};
    dim3 grid(ntx, nty, (g.kmax + t.kchunk - 1) / t.kchunk);
    const size_t smem = mom2_smem(sizeof(TF), t.kchunk, ty);
    CUtensorMap tu, tv, tw, te, tut, tvt, twt, tth;
    const int by = ty + 2 * T2_H;
    if (!make_field_tmap<TF>(&tu, a.u, g, T2_PX, by) || !make_field_tmap<TF>(&tv, a.v, g, T2_PX, by) ||
        !make_field_tmap<TF>(&tw, a.w, g, T2_PX, by) || !make_field_tmap<TF>(&te, a.evisc, g, T2_PX, by) ||
        !make_field_tmap<TF>(&tut, a.ut, g, T2_W + 2, ty) || !make_field_tmap<TF>(&tvt, a.vt, g, T2_W + 2, ty) ||
        !make_field_tmap<TF>(&twt, a.wt, g, T2_W + 2, ty) || !make_field_tmap<TF>(&tth, buoy ? (const void*)a.th : (const void*)a.u, g, T2_W + 2, ty))
    { c->err = "cuTensorMapEncodeTiled failed"; return MHH_E_CUDA; }
#define M2(S, B, Y) do { \
        static size_t attr_smem_dev[64] = {0}; size_t& attr_smem = attr_smem_dev[c->device & 63];   /* the attribute is per device */ \
        if (attr_smem < smem) { CUDA_TRY(c, cudaFuncSetAttribute(mom2_kernel<TF, S, B, Y>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; } \
        mom2_kernel<TF, S, B, Y><<<grid, 32 * Y, smem, c->stream>>>(tu, tv, tw, te, tut, tvt, twt, tth, t, g); } while (0)
#define M2Y(S, B) do { if (ty == 4) M2(S, B, 4); else if (ty == 8) M2(S, B, 8); else if (ty == 12) M2(S, B, 12); else M2(S, B, 6); } while (0)
    if (surface && buoy) M2Y(true, true);
    else if (surface) M2Y(true, false);
    else if (buoy) M2Y(false, true);
    else M2Y(false, false);
#undef M2Y
#undef M2
    KCHECKN(c, "mom2_kernel");
    return MHH_OK;
}

// warp-specialised variant (tile3_kernels.cuh): one CTA of (3+nsc)*ty+1 warps per SM; the first scalar rides along
template <typename TF>
int mom3_launch(Ctx<TF>* c, const MomArgs<TF>& a, const ScalArgs<TF>* sc, bool surface, bool buoy)
{
    const GridDev<TF>& g = c->g;
    const int nsc = sc ? 1 : 0;
    const int ty = c->tile3_y ? c->tile3_y : (nsc ? 3 : 4);
    const int ntx = (g.imax + T2_W - 1) / T2_W, nty = (g.jmax + ty - 1) / ty;
    Tend3Args<TF> t{};
    t.m = a; if (sc) t.sc = *sc;
    t.kchunk = pick_kchunk_waves(ntx * nty, g.kmax, c->num_sms, 2);
    t.prefetch = c->prefetch;
    dim3 grid(ntx, nty, (g.kmax + t.kchunk - 1) / t.kchunk);
    const size_t smem = mom3_smem(sizeof(TF), t.kchunk, ty, nsc);
    CUtensorMap tu, tv, tw, te, ts, tut, tvt, twt, tst;
    const int by = ty + 2 * T2_H;
    if (!make_field_tmap<TF>(&tu, a.u, g, T2_PX, by) || !make_field_tmap<TF>(&tv, a.v, g, T2_PX, by) ||
        !make_field_tmap<TF>(&tw, a.w, g, T2_PX, by) || !make_field_tmap<TF>(&te, a.evisc, g, T2_PX, by) ||
        !make_field_tmap<TF>(&ts, sc ? (const void*)sc->s : (const void*)a.u, g, T2_PX, by) ||
        !make_field_tmap<TF>(&tut, a.ut, g, T2_W + 2, ty) || !make_field_tmap<TF>(&tvt, a.vt, g, T2_W + 2, ty) ||
        !make_field_tmap<TF>(&twt, a.wt, g, T2_W + 2, ty) || !make_field_tmap<TF>(&tst, sc ? (const void*)sc->st : (const void*)a.ut, g, T2_W + 2, ty))
    { c->err = "cuTensorMapEncodeTiled failed"; return MHH_E_CUDA; }
#define M3(S, B, N, Y) do { \
        static size_t attr_smem_dev[64] = {0}; size_t& attr_smem = attr_smem_dev[c->device & 63];   /* the attribute is per device */ \
        if (attr_smem < smem) { CUDA_TRY(c, cudaFuncSetAttribute(mom3_kernel<TF, S, B, N, Y>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; } \
        mom3_kernel<TF, S, B, N, Y><<<grid, 32 * ((3 + N) * Y + 1), smem, c->stream>>>(tu, tv, tw, te, ts, tut, tvt, twt, tst, t, g); } while (0)
#define M3Y(S, B, N) do { if (ty == 3) M3(S, B, N, 3); else if (ty == 5) M3(S, B, N, 5); else M3(S, B, N, 4); } while (0)
    if (nsc)
    {
        if (surface && buoy) M3Y(true, true, 1);
        else if (surface) M3Y(true, false, 1);
        else if (buoy) M3Y(false, true, 1);
        else M3Y(false, false, 1);
    }
    else
    {
        if (surface && buoy) M3Y(true, true, 0);
        else if (surface) M3Y(true, false, 0);
        else if (buoy) M3Y(false, true, 0);
        else M3Y(false, false, 0);
    }
#undef M3Y
#undef M3
    KCHECKN(c, "mom3_kernel");
    return MHH_OK;
}

// 2 x 2 register-blocked warp-specialised variant (tile4_kernels.cuh); fp64 and fp32
template <typename TF>
int mom4_launch(Ctx<TF>* c, const MomArgs<TF>& a, const ScalArgs<TF>* sc, bool surface, bool buoy)
{
    const GridDev<TF>& g = c->g;
    const int nsc = sc ? 1 : 0;
    const int tyw = c->tile4_w;
    const int rows = 2 * tyw;
    const int ntx = (g.imax + T4_W - 1) / T4_W, nty = (g.jmax + rows - 1) / rows;
    Tend3Args<TF> t{};
    t.m = a; if (sc) t.sc = *sc;
    t.kchunk = pick_kchunk_waves(ntx * nty, g.kmax, c->num_sms, 2);
    t.prefetch = c->prefetch;
    dim3 grid(ntx, nty, (g.kmax + t.kchunk - 1) / t.kchunk);
    const size_t smem = mom4_smem(sizeof(TF), t.kchunk, tyw, nsc);
    CUtensorMap tu, tv, tw, te, ts, tut, tvt, twt, tst;
    const int by = t4_rows(tyw);
    constexpr int T4_PX = t4_px((int)sizeof(TF));
    if (!make_field_tmap<TF>(&tu, a.u, g, T4_PX, by) || !make_field_tmap<TF>(&tv, a.v, g, T4_PX, by) ||
        !make_field_tmap<TF>(&tw, a.w, g, T4_PX, by) || !make_field_tmap<TF>(&te, a.evisc, g, T4_PX, by) ||
        !make_field_tmap<TF>(&ts, sc ? (const void*)sc->s : (const void*)a.u, g, T4_PX, by) ||
        !make_field_tmap<TF>(&tut, a.ut, g, T4_PX, rows) || !make_field_tmap<TF>(&tvt, a.vt, g, T4_PX, rows) ||
        !make_field_tmap<TF>(&twt, a.wt, g, T4_PX, rows) || !make_field_tmap<TF>(&tst, sc ? (const void*)sc->st : (const void*)a.ut, g, T4_PX, rows))
    { c->err = "cuTensorMapEncodeTiled failed"; return MHH_E_CUDA; }
#define M4(S, B, N, W) do { \
        static size_t attr_smem_dev[64] = {0}; size_t& attr_smem = attr_smem_dev[c->device & 63];   /* the attribute is per device */ \
        if (attr_smem < smem) { CUDA_TRY(c, cudaFuncSetAttribute(mom4_kernel<TF, S, B, N, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; } \
        mom4_kernel<TF, S, B, N, W><<<grid, 32 * ((3 + N) * W + 1), smem, c->stream>>>(tu, tv, tw, te, ts, tut, tvt, twt, tst, t, g); } while (0)
#define M4W(S, B, N) do { if (tyw == 2) M4(S, B, N, 2); else M4(S, B, N, 3); } while (0)
    if (nsc)
    {
        if (surface && buoy) M4W(true, true, 1);
        else if (surface) M4W(true, false, 1);
        else if (buoy) M4W(false, true, 1);
        else M4W(false, false, 1);
    }
    else
    {
        if (surface && buoy) M4W(true, true, 0);
        else if (surface) M4W(true, false, 0);
        else if (buoy) M4W(false, true, 0);
        else M4W(false, false, 0);
    }
#undef M4W
#undef M4
    KCHECKN(c, "mom4_kernel");
    return MHH_OK;
}

template <typename TF>
int scal_tile_launch(Ctx<TF>* c, const ScalArgs<TF>& a, bool surface)
{
    const GridDev<TF>& g = c->g;
    const int ty = c->tile_y;
    const int ntx = (g.imax + TILE_X - 1) / TILE_X, nty = (g.jmax + ty - 1) / ty;
    ScalTileArgs<TF> t{a, pick_kchunk(ntx * nty, g.kmax, c->num_sms)};
    dim3 grid(ntx, nty, (g.kmax + t.kchunk - 1) / t.kchunk);
    const size_t smem = scal_tile_smem(sizeof(TF), t.kchunk, ty);
    const int vec = vec_width<TF>(g, {a.s, a.evisc});
#define ST3(S, V, Y) do { \
        static size_t attr_smem_dev[64] = {0}; size_t& attr_smem = attr_smem_dev[c->device & 63];   /* the attribute is per device */ \
        if (attr_smem < smem) { CUDA_TRY(c, cudaFuncSetAttribute(scal_tile_kernel<TF, S, V, Y>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; } \
        scal_tile_kernel<TF, S, V, Y><<<grid, TILE_X * Y, smem, c->stream>>>(t, g); } while (0)
#define ST(S, V) do { if (ty == 16) ST3(S, V, 16); else ST3(S, V, 8); } while (0)
    if (surface) { if (vec == 2) ST(true, 2); else ST(true, 1); }
    else { if (vec == 2) ST(false, 2); else ST(false, 1); }
#undef ST
#undef ST3
    KCHECKN(c, "scal_tile_kernel");
    return MHH_OK;
}

// tendencies: adv / diff / buoyancy in any combination (templates keep the unused parts out)
template <typename TF>
int tend_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, bool adv, bool diff, bool buoy)
{
    NEED_BASE(c);
    const GridDev<TF>& g = c->g;
    const bool surface = diff && prm && prm->surface_model;
    int rc = check_mom<TF>(c, f, diff, surface);
    if (rc != MHH_OK) return rc;
    if (adv && (g.igc < 3 || g.jgc < 3)) { c->err = "advec_2i5 needs igc, jgc >= 3"; return MHH_E_INVALID; }
    if (buoy && (f->ns < 1)) { c->err = "buoyancy needs scalar 0 (th)"; return MHH_E_INVALID; }
    const MomArgs<TF> a = mom_args<TF>(f);
    dim3 gr = c->grd_interior(), b = c->blk();
#define LAUNCH_MOM(A, D, S, B) tend_uvw_kernel<TF, A, D, S, B><<<gr, b, 0, c->stream>>>(a, g)
    const bool tiles = adv && diff && g.igc >= TILE_H && g.jgc >= TILE_H && !c->force_plain;
    int first_scalar = 0;       // scalars [0, first_scalar) were handled by the fused momentum kernel
    if (tiles)
    {
        const bool tma_any = !c->no_tma && tma_ok<TF>(g, {a.u, a.v, a.w, a.evisc, a.ut, a.vt, a.wt, buoy ? (const void*)a.th : (const void*)a.u});
        const bool tma = tma_any && sizeof(TF) == 8;          // mom2 / mom3 are fp64-only (odd-aligned 8-byte pairs)
        // the TMA box origin istart - halo must be 16-byte aligned (fp64: igc odd, fp32: igc a multiple of 4)
        const bool origin_ok = ((g.igc - t4_hl((int)sizeof(TF))) * (int)sizeof(TF)) % 16 == 0;
        if (tma_any && origin_ok && c->mom_variant == 4)
        {
            // fp64 and fp32 (USESP: needs a row pitch that is a multiple of 16 bytes, i.e. icells % 4 == 0 -- igc = 4)
            ScalArgs<TF> s0{};
            bool fuse = f->ns > 0 && c->fuse_scalar && !f->s_fluxlimit[0];
            if (fuse) { s0 = scal_args<TF>(c, f, prm, 0); fuse = tma_ok<TF>(g, {s0.s, s0.st}); }
            rc = mom4_launch<TF>(c, a, fuse ? &s0 : nullptr, surface, buoy);
            if (fuse) first_scalar = 1;
        }
        else if (tma && c->mom_variant == 3)
        {
            // scalar 0 rides along as the fourth warp group when its arrays qualify for TMA too
            ScalArgs<TF> s0{};
            // measured on B200 fp64 (512^3): 5.8 ms fused (3 rows, 13 warps) vs 4.3 + 2.9 ms as two kernels
            bool fuse = f->ns > 0 && c->fuse_scalar && !f->s_fluxlimit[0];
            if (fuse) { s0 = scal_args<TF>(c, f, prm, 0); fuse = tma_ok<TF>(g, {s0.s, s0.st}); }
            rc = mom3_launch<TF>(c, a, fuse ? &s0 : nullptr, surface, buoy);
            if (fuse) first_scalar = 1;
        }
        else if (tma) rc = mom2_launch<TF>(c, a, surface, buoy);
        else rc = mom_tile_launch<TF>(c, a, surface, buoy);
        if (rc != MHH_OK) return rc;
    }
    else if (adv && diff && surface && buoy) LAUNCH_MOM(true, true, true, true);
    else if (adv && diff && surface) LAUNCH_MOM(true, true, true, false);
    else if (adv && diff && buoy) LAUNCH_MOM(true, true, false, true);
    else if (adv && diff) LAUNCH_MOM(true, true, false, false);
    else if (adv) LAUNCH_MOM(true, false, false, false);
    else if (diff && surface) LAUNCH_MOM(false, true, true, false);
    else if (diff) LAUNCH_MOM(false, true, false, false);
    else { c->err = "tend_impl: nothing to do"; return MHH_E_INVALID; }
#undef LAUNCH_MOM
    if (!tiles) KCHECKN(c, "tend_uvw_kernel");
    for (int n = first_scalar; n < f->ns; ++n)
    {
        const ScalArgs<TF> s = scal_args<TF>(c, f, prm, n);
        if (adv && f->s_fluxlimit[n])
        {
            // `fluxlimit_list` scalar (src/advec_2i5.cxx:1046-1056): Koren-limited advection, then the diffusion alone
            advec_s_lim_kernel<TF><<<gr, b, 0, c->stream>>>(s.st, s.s, s.u, s.v, s.w, g);
            KCHECKN(c, "advec_s_lim_kernel");
            if (diff)
            {
                if (surface) tend_s_kernel<TF, false, true, true><<<gr, b, 0, c->stream>>>(s, g);
                else tend_s_kernel<TF, false, true, false><<<gr, b, 0, c->stream>>>(s, g);
                KCHECKN(c, "tend_s_kernel");
            }
            continue;
        }
        if (tiles)
        {
            if ((rc = scal_tile_launch<TF>(c, s, surface)) != MHH_OK) return rc;
            continue;
        }
#define LAUNCH_S(A, D, S) tend_s_kernel<TF, A, D, S><<<gr, b, 0, c->stream>>>(s, g)
        if (adv && diff && surface) LAUNCH_S(true, true, true);
        else if (adv && diff) LAUNCH_S(true, true, false);
        else if (adv) LAUNCH_S(true, false, false);
        else if (diff && surface) LAUNCH_S(false, true, true);
        else LAUNCH_S(false, true, false);
#undef LAUNCH_S
        KCHECKN(c, "tend_s_kernel");
    }
    return MHH_OK;
}

// Advec_2 / Diff_2 / thermo_dry buoyancy in any combination (order2_kernels.cuh)
template <typename TF>
int o2_impl(Ctx<TF>* c, const mhh_fields* f, bool adv, bool diff, bool buoy)
{
    NEED_BASE(c);
    const GridDev<TF>& g = c->g;
    int rc = check_mom<TF>(c, f, false, false);
    if (rc != MHH_OK) return rc;
    if (buoy && f->ns < 1) { c->err = "buoyancy needs scalar 0 (th)"; return MHH_E_INVALID; }
    if (!adv && !diff && !buoy) { c->err = "o2_impl: nothing to do"; return MHH_E_INVALID; }
    O2Args<TF> a{};
    a.ut = P<TF>(f->ut); a.vt = P<TF>(f->vt); a.wt = P<TF>(f->wt);
    a.u = P<TF>(f->u); a.v = P<TF>(f->v); a.w = P<TF>(f->w);
    a.th = f->ns > 0 ? P<TF>(f->s[0]) : nullptr;
    a.visc = (TF)f->visc;
    // `const double dxidxi = 1/(dx*dx);` (src/diff_2.cxx:44-45): the division itself is done in TF, then widened
    a.dxidxi = (double)(TF(1) / (g.dx * g.dx)); a.dyidyi = (double)(TF(1) / (g.dy * g.dy));
    dim3 gr = c->grd_interior(), b = c->blk();
#define O2(A, D, B) o2_uvw_kernel<TF, A, D, B><<<gr, b, 0, c->stream>>>(a, g)
    if (adv && diff && buoy) O2(true, true, true);
    else if (adv && diff) O2(true, true, false);
    else if (adv && buoy) O2(true, false, true);
    else if (adv) O2(true, false, false);
    else if (diff && buoy) O2(false, true, true);
    else if (diff) O2(false, true, false);
    else O2(false, false, true);
#undef O2
    KCHECKN(c, "o2_uvw_kernel");
    if (!adv && !diff) return MHH_OK;
    for (int n = 0; n < f->ns; ++n)
    {
        O2ScalArgs<TF> s{P<TF>(f->st[n]), P<TF>(f->s[n]), a.u, a.v, a.w, (TF)f->svisc[n], a.dxidxi, a.dyidyi};
        if (adv && diff) o2_s_kernel<TF, true, true><<<gr, b, 0, c->stream>>>(s, g);
        else if (adv) o2_s_kernel<TF, true, false><<<gr, b, 0, c->stream>>>(s, g);
        else o2_s_kernel<TF, false, true><<<gr, b, 0, c->stream>>>(s, g);
        KCHECKN(c, "o2_s_kernel");
    }
    return MHH_OK;
}

// Advec_4 / Diff_4 in any combination (order4_kernels.cuh)
template <typename TF>
int o4_impl(Ctx<TF>* c, const mhh_fields* f, bool adv, bool diff)
{
    const GridDev<TF>& g = c->g;
    if (!g.dzi4) { c->err = "4th-order schemes need a 4th-order grid (dzi4 / dzhi4 in mhh_grid_desc, three ghost cells)"; return MHH_E_INVALID; }
    if (g.kmax < 4) { c->err = "4th-order schemes need ktot >= 4"; return MHH_E_INVALID; }
    int rc = check_mom<TF>(c, f, false, false);
    if (rc != MHH_OK) return rc;
    O4Args<TF> a{};
    a.ut = P<TF>(f->ut); a.vt = P<TF>(f->vt); a.wt = P<TF>(f->wt);
    a.u = P<TF>(f->u); a.v = P<TF>(f->v); a.w = P<TF>(f->w);
    a.visc = (TF)f->visc;
    a.dxidxi_c = (TF)(1. / (double)(g.dx * g.dx)); a.dyidyi_c = (TF)(1. / (double)(g.dy * g.dy));
    a.dxidxi_w = TF(1) / (g.dx * g.dx); a.dyidyi_w = TF(1) / (g.dy * g.dy);
    const bool dim3 = g.jtot > 1;
    {
        ::dim3 gr = c->grd_interior(), b = c->blk();
#define O4(A, D, T) o4_uvw_kernel<TF, A, D, T><<<gr, b, 0, c->stream>>>(a, g)
        if (adv && diff) { if (dim3) O4(true, true, true); else O4(true, true, false); }
        else if (adv) { if (dim3) O4(true, false, true); else O4(true, false, false); }
        else if (diff) { if (dim3) O4(false, true, true); else O4(false, true, false); }
        else { c->err = "o4_impl: nothing to do"; return MHH_E_INVALID; }
#undef O4
        KCHECKN(c, "o4_uvw_kernel");
        for (int n = 0; n < f->ns; ++n)
        {
            O4ScalArgs<TF> s{P<TF>(f->st[n]), P<TF>(f->s[n]), a.u, a.v, a.w, (TF)f->svisc[n], a.dxidxi_c, a.dyidyi_c};
#define O4S(A, D, T) o4_s_kernel<TF, A, D, T><<<gr, b, 0, c->stream>>>(s, g)
            if (adv && diff) { if (dim3) O4S(true, true, true); else O4S(true, true, false); }
            else if (adv) { if (dim3) O4S(true, false, true); else O4S(true, false, false); }
            else { if (dim3) O4S(false, true, true); else O4S(false, true, false); }
#undef O4S
            KCHECKN(c, "o4_s_kernel");
        }
    }
    return MHH_OK;
}

template <typename TF>
int o2_cfl_impl(Ctx<TF>* c, const mhh_fields* f, double* out, int order = 2)
{
    const GridDev<TF>& g = c->g;
    NEED(c, f->u, "u"); NEED(c, f->v, "v"); NEED(c, f->w, "w");
    CUDA_TRY(c, cudaMemsetAsync(c->d_red, 0, sizeof(double), c->stream));
    if (order == 4) o4_cfl_kernel<TF><<<c->grd_interior(), c->blk(), 0, c->stream>>>(P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), g, c->d_red);
    else o2_cfl_kernel<TF><<<c->grd_interior(), c->blk(), 0, c->stream>>>(P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), g, c->d_red);
    const char* cfl_name = order == 4 ? "o4_cfl_kernel" : "o2_cfl_kernel";
    KCHECKN(c, cfl_name);
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

template <typename TF, int MODE>
int reduce_impl(Ctx<TF>* c, const TF* u, const TF* v, const TF* w, TF p0, TF p1, TF p2, double* out)
{
    const GridDev<TF>& g = c->g;
    CUDA_TRY(c, cudaMemsetAsync(c->d_red, 0, sizeof(double), c->stream));
    reduce_kernel<TF, MODE><<<c->grd_interior(), c->blk(), 0, c->stream>>>(u, v, w, g, p0, p1, p2, c->d_red);
    KCHECKN(c, "reduce_kernel");
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

// ---- Pres_2 ---------------------------------------------------------------------------------
// The slab all-to-all (reference semantics: Transpose::exec_xy / exec_yx, src/transpose.cxx:117-271).  Block d of the
// x-side buffer IS the message for rank d and block s of the y-side buffer IS the message from rank s (SpecLayout), so
// there is no pack/unpack pass: grouped ncclSend/ncclRecv straight out of / into the workspaces.
// All ranks have finished the kernels enqueued before this point once the all-reduce completes (it cannot finish
// before every rank has contributed, and every rank contributes in stream order after its own kernels).
template <typename TF>
int slab_barrier(Ctx<TF>* c, const char* name)
{
    NcclApi* api = nccl_api(c->err);
    if (!api) return MHH_E_CUDA;
    NCCL_TRY(c, api, api->AllReduce(c->d_barrier, c->d_barrier, 1, ncclInt32, ncclMax, c->comm, c->stream));
    prof_mark(c, name);
    return MHH_OK;
}

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

template <typename TF>
int pres_spectral_solve(Ctx<TF>* c, bool do_solve)
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
    if (do_solve)
    {
        const long long ncol = (long long)mcl * g.jtot;
        tdma_solve_kernel<TF><<<(unsigned)((ncol + 127) / 128), 128, 0, c->stream>>>(c->specT, c->fac, c->coef(), c->lay, mcl, g.jtot, g.kmax, c->lay.m_off, 0);
        KCHECKN(c, "tdma_solve_kernel");
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
    RhsSrc<TF> src{P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), P<TF>(f->ut), P<TF>(f->vt), P<TF>(f->wt), (TF)(TF(1.) / sub_dt), slab ? 0 : 1};
    const long long nrows = (long long)g.jmax * g.ktot;
    const int grid_x = (int)std::min<long long>((nrows + c->rows_x - 1) / c->rows_x, (long long)c->num_sms * 4);
    const int grid_wx = (int)std::min<long long>((nrows + WFFT_WARPS - 1) / WFFT_WARPS, (long long)c->num_sms * 8);
    if (c->wfft_x) wfft_x_forward_launch<TF>(g.itot / 2, true, grid_wx, c->stream, c->spec, src, g, c->lay, c->peers, c->tw_xh, c->tw_xf, nrows);
    else fft_x_forward_kernel<TF, true><<<grid_x, 256, c->smem_x, c->stream>>>(c->spec, src, g, c->lay, c->peers, c->plan_x, c->tw_xh, c->tw_xf, c->rows_x, nrows);
    KCHECKN(c, "fft_x_forward_kernel");
    int rc = pres_spectral_solve<TF>(c, true);
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

template <typename TF>
int rk3_impl(Ctx<TF>* c, TF* a, TF* at, int substep, double dt)
{
    const GridDev<TF>& g = c->g;
    NEED(c, a, "field"); NEED(c, at, "tendency");
    if (substep < 0 || substep > 2) { c->err = "substep must be 0..2"; return MHH_E_INVALID; }
    const TF cA[3] = {TF(0.), TF(-5. / 9.), TF(-153. / 128.)};
    const TF cB[3] = {TF(1. / 3.), TF(15. / 16.), TF(8. / 15.)};
    const int nxt = (substep + 1) % 3;
    rk3_kernel<TF><<<c->grd_all(), c->blk(), 0, c->stream>>>(a, at, cB[substep] * (TF)dt, cA[nxt], nxt == 0, g);
    KCHECKN(c, "rk3_kernel");
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
    if (c->nranks > 1) { c->err = "Pres_4 is single-GPU in this version"; return MHH_E_INVALID; }
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
    const long long ncol = (long long)c->nm * g.jtot;
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
    hdma_setup_kernel<TF><<<(unsigned)((ncol + 127) / 128), 128, 0, c->stream>>>(c->lu4, cf, c->nm, g.jtot, kmax);
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
    const TF dti = (TF)(1. / sub_dt);
    const long long pitch = 2 * c->nm;
    if (dim3) pres4_in_kernel<TF, true><<<c->grd_interior(), c->blk(), 0, c->stream>>>(P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), P<TF>(f->ut), P<TF>(f->vt), P<TF>(f->wt), c->spec, pitch, dti, g);
    else pres4_in_kernel<TF, false><<<c->grd_interior(), c->blk(), 0, c->stream>>>(P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), P<TF>(f->ut), P<TF>(f->vt), P<TF>(f->wt), c->spec, pitch, dti, g);
    KCHECKN(c, "pres4_in_kernel");
    const long long nrows = (long long)g.jmax * g.ktot;
    const int grid_x = (int)std::min<long long>((nrows + c->rows_x - 1) / c->rows_x, (long long)c->num_sms * 4);
    const int grid_wx = (int)std::min<long long>((nrows + WFFT_WARPS - 1) / WFFT_WARPS, (long long)c->num_sms * 8);
    RhsSrc<TF> none{};
    if (c->wfft_x) wfft_x_forward_launch<TF>(g.itot / 2, false, grid_wx, c->stream, c->spec, none, g, c->lay, c->peers, c->tw_xh, c->tw_xf, nrows);
    else fft_x_forward_kernel<TF, false><<<grid_x, 256, c->smem_x, c->stream>>>(c->spec, none, g, c->lay, c->peers, c->plan_x, c->tw_xh, c->tw_xf, c->rows_x, nrows);
    KCHECKN(c, "fft_x_forward_kernel");
    const int grid_p = c->num_sms * 2;
    const long long ypanels = (long long)((c->nm + WFFT_WARPS - 1) / WFFT_WARPS) * g.ktot;
    const int grid_wy = (int)std::max<long long>(1, std::min<long long>(ypanels, (long long)c->num_sms * 8));
    if (dim3)
    {
        if (c->wfft_y) wfft_y_launch<TF>(g.jtot, grid_wy, c->stream, c->spec, c->lay, c->peers, c->nm, g.ktot, c->tw_y, 0);
        else fft_y_kernel<TF><<<grid_p, 256, c->smem_y, c->stream>>>(c->spec, c->lay, c->peers, c->nm, g.jtot, g.ktot, c->plan_y, c->tw_y, c->mc_y, 0);
        KCHECKN(c, "fft_y_forward_kernel");
    }
    const long long ncol = (long long)c->nm * g.jtot;
    hdma_solve_kernel<TF><<<(unsigned)((ncol + 127) / 128), 128, 0, c->stream>>>(c->spec, c->lu4, c->nm, g.jtot, g.kmax);
    KCHECKN(c, "hdma_solve_kernel");
    if (dim3)
    {
        if (c->wfft_y) wfft_y_launch<TF>(g.jtot, grid_wy, c->stream, c->spec, c->lay, c->peers, c->nm, g.ktot, c->tw_y, 1);
        else fft_y_kernel<TF><<<grid_p, 256, c->smem_y, c->stream>>>(c->spec, c->lay, c->peers, c->nm, g.jtot, g.ktot, c->plan_y, c->tw_y, c->mc_y, 1);
        KCHECKN(c, "fft_y_backward_kernel");
    }
    const TF norm = TF(1.) / ((TF)g.itot * (TF)g.jtot);
    if (c->wfft_x) wfft_x_backward_launch<TF>(g.itot / 2, grid_wx, c->stream, c->spec, P<TF>(f->p), g, c->lay, c->tw_xh, c->tw_xf, nrows, norm, 1);
    else fft_x_backward_kernel<TF><<<grid_x, 256, c->smem_x, c->stream>>>(c->spec, P<TF>(f->p), g, c->lay, c->plan_x, c->tw_xh, c->tw_xf, c->rows_x, nrows, norm, 1);
    KCHECKN(c, "fft_x_backward_kernel");
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
    CUDA_TRY(c, cudaMemcpyAsync(c->h_red, c->d_red, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *out = *c->h_red;
    return MHH_OK;
}

// Boundary::set_ghost_cells, 4th order, one field (src/boundary.cxx:776-848, 963-991)
template <typename TF>
int ghost4_impl(Ctx<TF>* c, TF* fld, int bcbot, const TF* bot, const TF* gradbot, int bctop, const TF* top, const TF* gradtop)
{
    const GridDev<TF>& g = c->g;
    NEED(c, fld, "field");
    if (!g.dzi4) { c->err = "4th-order ghost cells need a 4th-order grid"; return MHH_E_INVALID; }
    if (bcbot == MHH_BC_DIRICHLET) NEED(c, bot, "bot");
    if (bcbot == MHH_BC_NEUMANN) NEED(c, gradbot, "gradbot");
    if (bctop == MHH_BC_DIRICHLET) NEED(c, top, "top");
    if (bctop == MHH_BC_NEUMANN) NEED(c, gradtop, "gradtop");
    // grad4(a,b,c,d) = -cg0*(d-a) - cg1*(c-b) (include/finite_difference.h:127-131) of the z levels around the walls
    const std::vector<TF>& z = c->h_z;
    auto grad4 = [](TF a, TF b, TF cc, TF d) { return -W4<TF>::cg0 * (d - a) - W4<TF>::cg1 * (cc - b); };
    const TF gb = grad4(z[g.kstart - 2], z[g.kstart - 1], z[g.kstart], z[g.kstart + 1]);
    const TF gt = grad4(z[g.kend - 2], z[g.kend - 1], z[g.kend], z[g.kend + 1]);
    dim3 b(64, 4), gr((g.icells + 63) / 64, (g.jcells + 3) / 4);
    ghost_cells_4th_kernel<TF><<<gr, b, 0, c->stream>>>(fld, g, bcbot, bot, gradbot, bctop, top, gradtop, gb, gt);
    KCHECKN(c, "ghost_cells_4th_kernel");
    return MHH_OK;
}

template <typename TF>
int ghost4w_impl(Ctx<TF>* c, TF* w, int conservation)
{
    const GridDev<TF>& g = c->g;
    NEED(c, w, "w");
    if (!g.dzi4) { c->err = "4th-order ghost cells need a 4th-order grid"; return MHH_E_INVALID; }
    dim3 b(64, 4), gr((g.icells + 63) / 64, (g.jcells + 3) / 4);
    ghost_cells_w_4th_kernel<TF><<<gr, b, 0, c->stream>>>(w, g, conservation);
    KCHECKN(c, "ghost_cells_w_4th_kernel");
    return MHH_OK;
}

// The 4th-order DNS sub-step (swspatialorder = 4: advec_4 + diff_4 + pres_4, no thermo), Model::exec order
// (src/model.cxx:368-437, 504): cyclic -> ghost cells (w normal) -> w conservation -> advec -> w normal -> diff ->
// w conservation -> pres -> w normal -> rk3.  advec_4 and diff_4 stay two launches: they see different w ghost cells
// (conservation vs normal type).
template <typename TF>
int substep_o4_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, int substep, double dt)
{
    if (c->nranks > 1) { c->err = "the 4th-order sub-step is single-GPU in this version"; return MHH_E_INVALID; }
    if (prm->swthermo != 0) { c->err = "the 4th-order sub-step has no thermo coupling (swthermo = 0)"; return MHH_E_INVALID; }
    int rc = check_mom<TF>(c, f, false, false);
    if (rc != MHH_OK) return rc;
    NEED(c, f->p, "p");
    TF* prog[3 + MHH_MAX_SCALARS] = {P<TF>(f->u), P<TF>(f->v), P<TF>(f->w)};
    for (int n = 0; n < f->ns; ++n) prog[3 + n] = P<TF>(f->s[n]);
    if ((rc = cyclic_fields<TF>(c, prog, 3 + f->ns)) != MHH_OK) return rc;
    if ((rc = ghost4_impl<TF>(c, P<TF>(f->u), prm->mbcbot, P<TF>(f->u_bot), P<TF>(f->u_gradbot), prm->mbctop, P<TF>(f->u_top), P<TF>(f->u_gradtop))) != MHH_OK) return rc;
    if ((rc = ghost4_impl<TF>(c, P<TF>(f->v), prm->mbcbot, P<TF>(f->v_bot), P<TF>(f->v_gradbot), prm->mbctop, P<TF>(f->v_top), P<TF>(f->v_gradtop))) != MHH_OK) return rc;
    for (int n = 0; n < f->ns; ++n)
        if ((rc = ghost4_impl<TF>(c, P<TF>(f->s[n]), prm->sbcbot[n], P<TF>(f->s_bot[n]), P<TF>(f->s_gradbot[n]),
                                  prm->sbctop[n], P<TF>(f->s_top[n]), P<TF>(f->s_gradtop[n]))) != MHH_OK) return rc;
    if ((rc = ghost4w_impl<TF>(c, P<TF>(f->w), 1)) != MHH_OK) return rc;        // (the normal-type fill right before is overwritten)
    if ((rc = o4_impl<TF>(c, f, true, false)) != MHH_OK) return rc;
    if ((rc = ghost4w_impl<TF>(c, P<TF>(f->w), 0)) != MHH_OK) return rc;
    if ((rc = o4_impl<TF>(c, f, false, true)) != MHH_OK) return rc;
    if ((rc = ghost4w_impl<TF>(c, P<TF>(f->w), 1)) != MHH_OK) return rc;
    const double cBd[3] = {1. / 3., 15. / 16., 8. / 15.};
    if ((rc = pres4_exec_impl<TF>(c, f, cBd[substep] * dt)) != MHH_OK) return rc;
    if ((rc = ghost4w_impl<TF>(c, P<TF>(f->w), 0)) != MHH_OK) return rc;
    TF* tend[3 + MHH_MAX_SCALARS] = {P<TF>(f->ut), P<TF>(f->vt), P<TF>(f->wt)};
    for (int n = 0; n < f->ns; ++n) tend[3 + n] = P<TF>(f->st[n]);
    for (int n = 0; n < 3 + f->ns; ++n)
        if ((rc = rk3_impl<TF>(c, prog[n], tend[n], substep, dt)) != MHH_OK) return rc;
    return MHH_OK;
}

// One fused sub-step (Model::exec order, src/model.cxx:356-504, restricted to the hot path), in three stages so that a host
// that keeps its own surface model can run it where the reference does (src/model.cxx:375-401: exec_viscosity ->
// thermo.exec -> boundary.exec + set_ghost_cells -> advec.exec ...):
//   pre  : boundary.set_prognostic_cyclic_bcs + set_ghost_cells, diff.exec_viscosity
//   ghost: boundary.set_ghost_cells again (after the host's boundary.exec changed the 2-D companions)
//   post : thermo.exec + advec.exec + diff.exec (fused), pres.exec, timeloop.exec
template <typename TF>
int substep_check(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, bool& o4)
{
    NEED_BASE(c);
    NEED(c, prm, "params");
    o4 = false;
    if (prm->swadvec == 4 || prm->swdiff == 4)
    {
        if (prm->swadvec != 4 || prm->swdiff != 4) { c->err = "dycore_substep: the 4th-order configuration is swadvec = 4 with swdiff = 4 (and pres_4)"; return MHH_E_INVALID; }
        o4 = true;
        return MHH_OK;
    }
    if ((prm->swadvec != 25 && prm->swadvec != 2) || (prm->swdiff != 1 && prm->swdiff != 2))
    { c->err = "dycore_substep: swadvec must be 2i5 (25), 2 or 4, swdiff smag2 (1), 2 or 4"; return MHH_E_INVALID; }
    const bool smag = prm->swdiff == 1;
    int rc = check_mom<TF>(c, f, smag, smag && prm->surface_model != 0);
    if (rc != MHH_OK) return rc;
    NEED(c, f->p, "p");
    return MHH_OK;
}

// Boundary::set_ghost_cells of u, v and the scalars (2nd order)
template <typename TF>
int ghost_all_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm)
{
    int rc;
    if ((rc = ghost_impl<TF>(c, P<TF>(f->u), prm->mbcbot, P<TF>(f->u_bot), P<TF>(f->u_gradbot), prm->mbctop, P<TF>(f->u_top), P<TF>(f->u_gradtop))) != MHH_OK) return rc;
    if ((rc = ghost_impl<TF>(c, P<TF>(f->v), prm->mbcbot, P<TF>(f->v_bot), P<TF>(f->v_gradbot), prm->mbctop, P<TF>(f->v_top), P<TF>(f->v_gradtop))) != MHH_OK) return rc;
    for (int n = 0; n < f->ns; ++n)
        if ((rc = ghost_impl<TF>(c, P<TF>(f->s[n]), prm->sbcbot[n], P<TF>(f->s_bot[n]), P<TF>(f->s_gradbot[n]),
                                 prm->sbctop[n], P<TF>(f->s_top[n]), P<TF>(f->s_gradtop[n]))) != MHH_OK) return rc;
    return MHH_OK;
}

template <typename TF>
int substep_pre_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm)
{
    bool o4;
    int rc = substep_check<TF>(c, f, prm, o4);
    if (rc != MHH_OK) return rc;
    if (o4) { c->err = "dycore_substep_pre/post: the 4th-order configuration has no surface model; use mhh_dycore_substep"; return MHH_E_INVALID; }
    // 1. boundary.set_prognostic_cyclic_bcs + set_ghost_cells
    TF* prog[3 + MHH_MAX_SCALARS] = {P<TF>(f->u), P<TF>(f->v), P<TF>(f->w)};
    for (int n = 0; n < f->ns; ++n) prog[3 + n] = P<TF>(f->s[n]);
    if ((rc = cyclic_fields<TF>(c, prog, 3 + f->ns)) != MHH_OK) return rc;
    if ((rc = ghost_all_impl<TF>(c, f, prm)) != MHH_OK) return rc;
    // 2. diff.exec_viscosity
    if (prm->swdiff == 1 && (rc = evisc_impl<TF>(c, f, prm, nullptr)) != MHH_OK) return rc;
    return MHH_OK;
}

// 3. thermo.exec + advec.exec + diff.exec: one fused kernel per scheme family
template <typename TF>
int tendencies_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm)
{
    bool o4;
    int rc = substep_check<TF>(c, f, prm, o4);
    if (rc != MHH_OK) return rc;
    if (o4) { c->err = "dycore_tendencies: use mhh_advec_exec / mhh_diff_4_exec on a 4th-order grid (they see different w ghost cells)"; return MHH_E_INVALID; }
    const bool smag = prm->swdiff == 1, adv5 = prm->swadvec == 25, buoy = prm->swthermo == 1;
    if (adv5 && smag) rc = tend_impl<TF>(c, f, prm, true, true, buoy);                       // 2i5 + smag2 (+ buoyancy)
    else if (!adv5 && !smag) rc = o2_impl<TF>(c, f, true, true, buoy);                       // 2 + 2 (+ buoyancy)
    else if (!adv5)                                                                          // 2 + smag2 (drycblles as shipped)
    {
        if ((rc = o2_impl<TF>(c, f, true, false, buoy)) != MHH_OK) return rc;
        rc = tend_impl<TF>(c, f, prm, false, true, false);
    }
    else                                                                                     // 2i5 + 2
    {
        if (buoy && (rc = o2_impl<TF>(c, f, false, false, true)) != MHH_OK) return rc;
        if ((rc = tend_impl<TF>(c, f, prm, true, false, false)) != MHH_OK) return rc;
        rc = o2_impl<TF>(c, f, false, true, false);
    }
    return rc;
}

template <typename TF>
int substep_post_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, int substep, double dt)
{
    const GridDev<TF>& g = c->g;
    if (substep < 0 || substep > 2) { c->err = "substep must be 0..2"; return MHH_E_INVALID; }
    int rc = tendencies_impl<TF>(c, f, prm);
    if (rc != MHH_OK) return rc;
    // 4. pres.exec (solve), then pressure correction fused with timeloop.exec
    const TF cA[3] = {TF(0.), TF(-5. / 9.), TF(-153. / 128.)};
    const TF cB[3] = {TF(1. / 3.), TF(15. / 16.), TF(8. / 15.)};
    const double cBd[3] = {1. / 3., 15. / 16., 8. / 15.};
    const double sub_dt = cBd[substep] * dt;          // Timeloop::get_sub_time_step (double)
    if ((rc = pres_solve_impl<TF>(c, f, sub_dt)) != MHH_OK) return rc;
    const int nxt = (substep + 1) % 3;
    PresArgs<TF> a{P<TF>(f->ut), P<TF>(f->vt), P<TF>(f->wt), P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), P<TF>(f->p)};
    pres_out_rk3_kernel<TF><<<c->grd_all(), c->blk(), 0, c->stream>>>(a, P<TF>(f->u), P<TF>(f->v), P<TF>(f->w),
            cB[substep] * (TF)dt, cA[nxt], nxt == 0, g);
    KCHECKN(c, "pres_out_rk3_kernel");
    for (int n = 0; n < f->ns; ++n)
        if ((rc = rk3_impl<TF>(c, P<TF>(f->s[n]), P<TF>(f->st[n]), substep, dt)) != MHH_OK) return rc;
    return MHH_OK;
}

template <typename TF>
int substep_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, int substep, double dt)
{
    bool o4;
    int rc = substep_check<TF>(c, f, prm, o4);
    if (rc != MHH_OK) return rc;
    if (substep < 0 || substep > 2) { c->err = "substep must be 0..2"; return MHH_E_INVALID; }
    if (o4) return substep_o4_impl<TF>(c, f, prm, substep, dt);
    if ((rc = substep_pre_impl<TF>(c, f, prm)) != MHH_OK) return rc;
    return substep_post_impl<TF>(c, f, prm, substep, dt);
}

template <typename TF>
int step_host_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, double dt, int nsteps,
                   void* h_u, void* h_v, void* h_w, void* const* h_s)
{
    const GridDev<TF>& g = c->g;
    const size_t bytes = sizeof(TF) * (size_t)g.ncells;
    NEED(c, h_u, "h_u"); NEED(c, h_v, "h_v"); NEED(c, h_w, "h_w");
    CUDA_TRY(c, cudaMemcpyAsync(f->u, h_u, bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(f->v, h_v, bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(f->w, h_w, bytes, cudaMemcpyHostToDevice, c->stream));
    for (int n = 0; n < f->ns; ++n)
    {
        NEED(c, h_s[n], "h_s[n]");
        CUDA_TRY(c, cudaMemcpyAsync(f->s[n], h_s[n], bytes, cudaMemcpyHostToDevice, c->stream));
    }
    for (int it = 0; it < nsteps; ++it)
        for (int ss = 0; ss < 3; ++ss)
        {
            int rc = substep_impl<TF>(c, f, prm, ss, dt);
            if (rc != MHH_OK) return rc;
        }
    CUDA_TRY(c, cudaMemcpyAsync(h_u, f->u, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(h_v, f->v, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(h_w, f->w, bytes, cudaMemcpyDeviceToHost, c->stream));
    for (int n = 0; n < f->ns; ++n)
        CUDA_TRY(c, cudaMemcpyAsync(h_s[n], f->s[n], bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return MHH_OK;
}

template <typename TF>
int fft_roundtrip_impl(Ctx<TF>* c, const TF* in, TF* out, int solve)
{
    NEED_BASE(c);
    const GridDev<TF>& g = c->g;
    NEED(c, in, "in"); NEED(c, out, "out");
    if (c->nranks > 1) { c->err = "pres_fft_roundtrip is a single-GPU test entry point"; return MHH_E_INVALID; }
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
    int rc = pres_spectral_solve<TF>(c, solve != 0);
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

} // namespace

// ============================================================================================
// C entry points
// ============================================================================================
#define DISPATCH(ctx, expr64, expr32) \
    do { if (!(ctx)) return MHH_E_INVALID; \
         cudaError_t e_ = cudaSetDevice((ctx)->device); \
         if (e_ != cudaSuccess) { (ctx)->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e_); return MHH_E_CUDA; } \
         if ((ctx)->dtype == MHH_F64) { typedef double TF; Ctx<TF>* c = static_cast<Ctx<TF>*>(ctx); (void)c; return (expr64); } \
         else { typedef float TF; Ctx<TF>* c = static_cast<Ctx<TF>*>(ctx); (void)c; return (expr32); } } while (0)
#define DISPATCH1(ctx, expr) DISPATCH(ctx, expr, expr)

extern "C" {

int mhh_ctx_create(const mhh_grid_desc* grid, int dtype, int device, mhh_ctx** out)
{
    if (!grid || !out) return MHH_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev)
    {
        // no CPU fallback: a context cannot exist without a CUDA device
        cudaGetLastError();
        return MHH_E_CUDA;
    }
    int rc;
    if (dtype == MHH_F64) rc = create_impl<double>(grid, dtype, device, out);
    else if (dtype == MHH_F32) rc = create_impl<float>(grid, dtype, device, out);
    else return MHH_E_INVALID;
    return rc;   // on failure *out stays valid so that mhh_last_error() can be read; caller destroys it
}

void mhh_ctx_destroy(mhh_ctx* ctx) { delete ctx; }

const char* mhh_last_error(const mhh_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context (no CUDA device?)"; }

int mhh_sync(mhh_ctx* ctx)
{
    if (!ctx) return MHH_E_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return MHH_OK;
}

int mhh_set_stream(mhh_ctx* ctx, void* s)
{
    if (!ctx) return MHH_E_INVALID;
    ctx->stream = static_cast<cudaStream_t>(s);   // NULL selects the CUDA legacy default stream
    return MHH_OK;
}

long long mhh_launch_count(const mhh_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mhh_comm_get_unique_id(void* id, int nbytes)
{
    if (!id || nbytes < (int)sizeof(ncclUniqueId)) return MHH_E_INVALID;
    std::string err;
    NcclApi* api = nccl_api(err);
    if (!api) return MHH_E_CUDA;
    ncclUniqueId u;
    if (api->GetUniqueId(&u) != ncclSuccess) return MHH_E_CUDA;
    memset(id, 0, (size_t)nbytes);
    memcpy(id, &u, sizeof(u));
    return MHH_OK;
}

int mhh_comm_init(mhh_ctx* ctx, const void* id, int nbytes)
{
    if (!ctx) return MHH_E_INVALID;
    if (!id || nbytes < (int)sizeof(ncclUniqueId)) { ctx->err = "comm_init: bad unique id"; return MHH_E_INVALID; }
    if (ctx->comm) { ctx->err = "comm_init: communicator already set"; return MHH_E_INVALID; }
    if (ctx->nranks == 1) return MHH_OK;              // nothing to connect
    NcclApi* api = nccl_api(ctx->err);
    if (!api) return MHH_E_CUDA;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    NCCL_TRY(ctx, api, api->CommInitRank(&ctx->comm, ctx->nranks, u, ctx->rank));
    return MHH_OK;
}

int mhh_comm_get_ipc_handles(mhh_ctx* ctx, void* out, int nbytes)
{
    if (!ctx) return MHH_E_INVALID;
    if (!out || nbytes < MHH_IPC_BYTES) { ctx->err = "get_ipc_handles: buffer too small"; return MHH_E_INVALID; }
    DISPATCH1(ctx, ([&]() -> int {
        cudaIpcMemHandle_t h[3];
        static_assert(3 * sizeof(cudaIpcMemHandle_t) == MHH_IPC_BYTES, "handle size");
        if (!c->phalo)
        {
            // receive buffers for the peer halos: up to 8 fields of jgc rows, two directions, two alternating sets
            const GridDev<TF>& g = c->g;
            c->phalo_cap = (size_t)8 * g.jgc * g.icells * g.kcells;
            CUDA_TRY(c, cudaMalloc(&c->phalo, sizeof(TF) * c->phalo_cap * 4));
            c->ws_bytes += (long long)(sizeof(TF) * c->phalo_cap * 4);
        }
        CUDA_TRY(c, cudaIpcGetMemHandle(&h[0], c->spec));
        CUDA_TRY(c, cudaIpcGetMemHandle(&h[1], c->specT));
        CUDA_TRY(c, cudaIpcGetMemHandle(&h[2], c->phalo));
        memcpy(out, h, sizeof(h));
        return MHH_OK; })());
}

int mhh_comm_open_peers(mhh_ctx* ctx, const void* all, int nbytes)
{
    if (!ctx) return MHH_E_INVALID;
    if (!all || nbytes < ctx->nranks * MHH_IPC_BYTES) { ctx->err = "open_peers: need nranks * MHH_IPC_BYTES bytes"; return MHH_E_INVALID; }
    if (ctx->nranks == 1) return MHH_OK;
    if (ctx->nranks > MAX_SLAB_RANKS) { ctx->err = "open_peers: too many ranks"; return MHH_E_INVALID; }
    if (!ctx->comm) { ctx->err = "open_peers: call mhh_comm_init first"; return MHH_E_INVALID; }
    { const char* e = getenv("MHH_NO_PEER"); if (e && e[0] == '1') return MHH_OK; }       // keep the NCCL all-to-all (A/B comparisons)
    { const char* e = getenv("MHH_FAIL_PEER_RANK"); if (e && atoi(e) == ctx->rank) { ctx->err = "open_peers: failure injected by MHH_FAIL_PEER_RANK (test knob)"; return MHH_E_CUDA; } }
    DISPATCH1(ctx, ([&]() -> int {
        if (c->peers.on) { c->err = "open_peers: already open"; return MHH_E_INVALID; }
        if (c->g.jtot == 1) return MHH_OK;
        const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(all);
        PeerPtrs<TF> pp{};
        for (int r = 0; r < c->nranks; ++r)
        {
            if (r == c->rank) { pp.x[r] = c->spec; pp.y[r] = c->specT; continue; }
            void *px = nullptr, *py = nullptr;
            CUDA_TRY(c, cudaIpcOpenMemHandle(&px, h[3 * r], cudaIpcMemLazyEnablePeerAccess));
            CUDA_TRY(c, cudaIpcOpenMemHandle(&py, h[3 * r + 1], cudaIpcMemLazyEnablePeerAccess));
            pp.x[r] = static_cast<TF*>(px); pp.y[r] = static_cast<TF*>(py);
        }
        if (c->phalo && !(getenv("MHH_NO_PEER_HALO") && getenv("MHH_NO_PEER_HALO")[0] == '1'))
        {
            const int south = (c->rank + c->nranks - 1) % c->nranks, north = (c->rank + 1) % c->nranks;
            void *ps = nullptr, *pn = nullptr;
            CUDA_TRY(c, cudaIpcOpenMemHandle(&ps, h[3 * south + 2], cudaIpcMemLazyEnablePeerAccess));
            if (north == south) pn = ps;
            else CUDA_TRY(c, cudaIpcOpenMemHandle(&pn, h[3 * north + 2], cudaIpcMemLazyEnablePeerAccess));
            c->phalo_south = static_cast<TF*>(ps); c->phalo_north = static_cast<TF*>(pn);
        }
        pp.on = 1;
        c->peers = pp;
        c->lay.xtiled = 1;
        return MHH_OK; })());
}

int mhh_comm_transport(const mhh_ctx* ctx)
{
    if (!ctx || ctx->nranks == 1 || !ctx->comm) return 0;
    if (ctx->dtype == MHH_F64) return static_cast<const Ctx<double>*>(ctx)->peers.on ? 2 : 1;
    return static_cast<const Ctx<float>*>(ctx)->peers.on ? 2 : 1;
}

int mhh_comm_disable_peers(mhh_ctx* ctx)
{
    if (!ctx) return MHH_E_INVALID;
    // back to grouped ncclSend/ncclRecv for transposes and ghost rows (collective decision of the host: every rank calls it)
    DISPATCH1(ctx, ([&]() -> int {
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        for (int r = 0; r < MAX_SLAB_RANKS; ++r)
            if (c->peers.on && r != c->rank)
            {
                if (c->peers.x[r]) cudaIpcCloseMemHandle(c->peers.x[r]);
                if (c->peers.y[r]) cudaIpcCloseMemHandle(c->peers.y[r]);
            }
        c->peers = PeerPtrs<TF>{};
        c->lay.xtiled = 0;
        if (c->phalo_south && c->phalo_south != c->phalo) cudaIpcCloseMemHandle(c->phalo_south);
        if (c->phalo_north && c->phalo_north != c->phalo && c->phalo_north != c->phalo_south) cudaIpcCloseMemHandle(c->phalo_north);
        c->phalo_south = nullptr; c->phalo_north = nullptr;
        cudaGetLastError();
        return MHH_OK; })());
}

int mhh_slab_layout(int itot, int jtot, int ktot, int npy, int rank, mhh_slab_info* out)
{
    if (!out || npy < 1 || rank < 0 || rank >= npy || itot < 2 || jtot < 1 || ktot < 1 || jtot % npy != 0 || itot / 2 + 1 < npy) return MHH_E_INVALID;
    const SpecLayout l = make_spec_layout(itot, jtot, ktot, npy, rank);
    out->nm = l.nm; out->mcl = l.mcl; out->m_off = l.m_off; out->jmax = l.jmax;
    out->rows = l.rows;
    out->xside_elems = (long long)l.nm * l.rows;
    out->yside_elems = (long long)l.mcl * jtot * ktot;
    return MHH_OK;
}

long long mhh_slab_xindex(int itot, int jtot, int ktot, int npy, int rank, long long row, int m)
{
    if (npy < 1 || rank < 0 || rank >= npy || jtot % npy != 0 || itot / 2 + 1 < npy) return -1;
    const SpecLayout l = make_spec_layout(itot, jtot, ktot, npy, rank);
    if (row < 0 || row >= l.rows || m < 0 || m >= l.nm) return -1;
    return l.xidx(row, m);
}

long long mhh_slab_xindex_tiled(int itot, int jtot, int ktot, int npy, int rank, long long row, int m, long long* total)
{
    if (npy < 1 || rank < 0 || rank >= npy || jtot % npy != 0 || itot / 2 + 1 < npy) return -1;
    SpecLayout l = make_spec_layout(itot, jtot, ktot, npy, rank);
    l.xtiled = 1;
    if (total) *total = l.xside_elems();
    if (row < 0 || row >= l.rows || m < 0 || m >= l.nm) return -1;
    return l.xidx(row, m);
}

long long mhh_slab_yindex(int itot, int jtot, int ktot, int npy, int rank, int k, int j, int ml)
{
    if (npy < 1 || rank < 0 || rank >= npy || jtot % npy != 0 || itot / 2 + 1 < npy) return -1;
    const SpecLayout l = make_spec_layout(itot, jtot, ktot, npy, rank);
    if (k < 0 || k >= ktot || j < 0 || j >= jtot || ml < 0 || ml >= l.mcl) return -1;
    return l.yidx(k, j, ml);
}

int mhh_profile_start(mhh_ctx* ctx)
{
    if (!ctx) return MHH_E_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    for (auto& pe : ctx->prof_events) ctx->prof_pool.push_back(pe.second);
    ctx->prof_events.clear();
    ctx->prof = true;
    prof_mark(ctx, "__start__");
    return MHH_OK;
}

int mhh_profile_stop(mhh_ctx* ctx, const char** json)
{
    if (!ctx || !json) return MHH_E_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->prof = false;
    std::vector<std::string> names; std::vector<double> ms; std::vector<long long> cnt;
    for (size_t n = 1; n < ctx->prof_events.size(); ++n)
    {
        float t = 0.f;
        cudaEventElapsedTime(&t, ctx->prof_events[n - 1].second, ctx->prof_events[n].second);
        const std::string nm = ctx->prof_events[n].first;
        size_t k = 0;
        for (; k < names.size(); ++k) if (names[k] == nm) break;
        if (k == names.size()) { names.push_back(nm); ms.push_back(0.); cnt.push_back(0); }
        ms[k] += t; cnt[k] += 1;
    }
    std::string js = "{";
    for (size_t k = 0; k < names.size(); ++k)
    {
        char buf[256];
        snprintf(buf, sizeof(buf), "%s\"%s\": {\"n\": %lld, \"ms\": %.6f}", k ? ", " : "", names[k].c_str(), cnt[k], ms[k]);
        js += buf;
    }
    js += "}";
    ctx->prof_json = js;
    *json = ctx->prof_json.c_str();
    for (auto& pe : ctx->prof_events) ctx->prof_pool.push_back(pe.second);
    ctx->prof_events.clear();
    return MHH_OK;
}
long long mhh_workspace_bytes(const mhh_ctx* ctx) { return ctx ? ctx->ws_bytes : 0; }

int mhh_set_basestate(mhh_ctx* ctx, const void* rhoref, const void* rhorefh, const void* thref, const void* threfh)
{ DISPATCH1(ctx, set_basestate_impl<TF>(c, rhoref, rhorefh, thref, threfh)); }

int mhh_boundary_cyclic(mhh_ctx* ctx, void* fld, int edge)
{ DISPATCH1(ctx, cyclic_impl<TF>(c, P<TF>(fld), edge, false)); }

int mhh_boundary_cyclic_2d(mhh_ctx* ctx, void* fld)
{ DISPATCH1(ctx, cyclic_impl<TF>(c, P<TF>(fld), MHH_EDGE_BOTH, true)); }

int mhh_boundary_ghost_cells_2nd(mhh_ctx* ctx, void* fld, int bcbot, const void* bot, const void* gradbot,
                                 int bctop, const void* top, const void* gradtop)
{ DISPATCH1(ctx, ghost_impl<TF>(c, P<TF>(fld), bcbot, P<TF>(bot), P<TF>(gradbot), bctop, P<TF>(top), P<TF>(gradtop))); }

int mhh_boundary_ghost_cells_4th(mhh_ctx* ctx, void* fld, int bcbot, const void* bot, const void* gradbot,
                                 int bctop, const void* top, const void* gradtop)
{ DISPATCH1(ctx, ghost4_impl<TF>(c, P<TF>(fld), bcbot, P<TF>(bot), P<TF>(gradbot), bctop, P<TF>(top), P<TF>(gradtop))); }

int mhh_boundary_ghost_cells_w_4th(mhh_ctx* ctx, void* w, int conservation)
{ DISPATCH1(ctx, ghost4w_impl<TF>(c, P<TF>(w), conservation)); }

int mhh_advec_exec(mhh_ctx* ctx, int swadvec, const mhh_fields* f)
{
    if (ctx && swadvec != 25 && swadvec != 2 && swadvec != 4) { ctx->err = "advec_exec: swadvec must be 25 (2i5), 2 or 4"; return MHH_E_INVALID; }
    if (swadvec == 2) DISPATCH1(ctx, o2_impl<TF>(c, f, true, false, false));
    if (swadvec == 4) DISPATCH1(ctx, o4_impl<TF>(c, f, true, false));
    DISPATCH1(ctx, tend_impl<TF>(c, f, nullptr, true, false, false));
}

int mhh_diff_2_exec(mhh_ctx* ctx, const mhh_fields* f)
{
    if (!f) return MHH_E_INVALID;
    DISPATCH1(ctx, o2_impl<TF>(c, f, false, true, false));
}

int mhh_diff_4_exec(mhh_ctx* ctx, const mhh_fields* f)
{
    if (!f) return MHH_E_INVALID;
    DISPATCH1(ctx, o4_impl<TF>(c, f, false, true));
}

int mhh_diff_2_get_dn(mhh_ctx* ctx, const mhh_fields* f, double dt, double* dn)
{
    if (!ctx || !f || !dn) return MHH_E_INVALID;
    if (f->ns < 0 || f->ns > MHH_MAX_SCALARS) { ctx->err = "ns out of range"; return MHH_E_INVALID; }
    // Diff_2::create + get_dn (src/diff_2.cxx:133-152): host arithmetic on the grid metrics, no field pass
    DISPATCH1(ctx, ([&]() -> int {
        double viscmax = (double)(TF)f->visc;
        for (int n = 0; n < f->ns; ++n) viscmax = std::max(viscmax, (double)(TF)f->svisc[n]);
        const GridDev<TF>& g = c->g;
        double dnmul = 0.;
        for (int k = g.kstart; k < g.kend; ++k)
        {
            const TF dzk = c->h_dz[k];
            dnmul = std::max(dnmul, std::abs((double)(TF)viscmax * (1. / (double)(g.dx * g.dx) + 1. / (double)(g.dy * g.dy) + 1. / (double)(dzk * dzk))));
        }
        *dn = dnmul * dt;
        return MHH_OK; })());
}

int mhh_advec_get_cfl(mhh_ctx* ctx, int swadvec, const mhh_fields* f, double dt, double* cfl)
{
    if (!ctx || !f || !cfl) return MHH_E_INVALID;
    if (swadvec != 25 && swadvec != 2 && swadvec != 4) { ctx->err = "advec_get_cfl: swadvec must be 25 (2i5), 2 or 4"; return MHH_E_INVALID; }
    int rc;
    if (swadvec == 2 || swadvec == 4)
    {
        if (ctx->dtype == MHH_F64) { rc = o2_cfl_impl<double>(static_cast<Ctx<double>*>(ctx), f, cfl, swadvec); if (rc == MHH_OK) *cfl = *cfl * dt; }
        else { rc = o2_cfl_impl<float>(static_cast<Ctx<float>*>(ctx), f, cfl, swadvec); if (rc == MHH_OK) *cfl = (double)((float)*cfl * (float)dt); }
        return rc;
    }
    if (ctx->dtype == MHH_F64) { typedef double TF; rc = reduce_impl<TF, 0>(static_cast<Ctx<TF>*>(ctx), P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), 0, 0, 0, cfl); if (rc == MHH_OK) *cfl = *cfl * dt; }
    else { typedef float TF; rc = reduce_impl<TF, 0>(static_cast<Ctx<TF>*>(ctx), P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), 0, 0, 0, cfl); if (rc == MHH_OK) *cfl = (double)((float)*cfl * (float)dt); }
    return rc;
}

int mhh_diff_smag2_exec_viscosity(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, const void* n2)
{
    if (!f || !prm) return MHH_E_INVALID;
    DISPATCH1(ctx, evisc_impl<TF>(c, f, prm, P<TF>(n2)));
}

int mhh_diff_smag2_exec(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm)
{
    if (!f || !prm) return MHH_E_INVALID;
    DISPATCH1(ctx, tend_impl<TF>(c, f, prm, false, true, false));
}

int mhh_diff_smag2_get_dn(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, double dt, double* dn)
{
    if (!ctx || !f || !prm || !dn) return MHH_E_INVALID;
    int rc;
    if (ctx->dtype == MHH_F64)
    {
        typedef double TF; Ctx<TF>* c = static_cast<Ctx<TF>*>(ctx);
        const TF tprfac = TF(1) / std::min(TF(1.), (TF)prm->tPr);
        rc = reduce_impl<TF, 1>(c, P<TF>(f->evisc), nullptr, nullptr, tprfac, (TF)(1. / ((double)c->g.dx * c->g.dx)), (TF)(1. / ((double)c->g.dy * c->g.dy)), dn);
    }
    else
    {
        typedef float TF; Ctx<TF>* c = static_cast<Ctx<TF>*>(ctx);
        const TF tprfac = TF(1) / std::min(TF(1.), (TF)prm->tPr);
        rc = reduce_impl<TF, 1>(c, P<TF>(f->evisc), nullptr, nullptr, tprfac, (TF)(1. / ((double)c->g.dx * c->g.dx)), (TF)(1. / ((double)c->g.dy * c->g.dy)), dn);
    }
    if (rc == MHH_OK) *dn = *dn * dt;
    return rc;
}

int mhh_thermo_dry_exec(mhh_ctx* ctx, void* wt, const void* th)
{
    if (!ctx || !wt || !th) return MHH_E_INVALID;
    DISPATCH1(ctx, ([&]() -> int {
        NEED_BASE(c);
        dim3 gr = c->grd_interior(); gr.z = c->g.kmax - 1;
        buoyancy_kernel<TF><<<gr, c->blk(), 0, c->stream>>>(P<TF>(wt), P<TF>(th), c->g);
        KCHECKN(c, "buoyancy_kernel"); return MHH_OK; })());
}

int mhh_thermo_dry_n2(mhh_ctx* ctx, void* n2, const void* th)
{
    if (!ctx || !n2 || !th) return MHH_E_INVALID;
    DISPATCH1(ctx, ([&]() -> int {
        NEED_BASE(c);
        n2_kernel<TF><<<c->grd_interior(), c->blk(), 0, c->stream>>>(P<TF>(n2), P<TF>(th), c->g);
        KCHECKN(c, "n2_kernel"); return MHH_OK; })());
}

int mhh_pres_exec(mhh_ctx* ctx, int swpres, const mhh_fields* f, double sub_dt)
{
    if (!f) return MHH_E_INVALID;
    if (ctx && swpres != 2 && swpres != 4) { ctx->err = "pres_exec: swpres must be 2 or 4"; return MHH_E_INVALID; }
    if (swpres == 4) DISPATCH1(ctx, pres4_exec_impl<TF>(c, f, sub_dt));
    DISPATCH1(ctx, pres_exec_impl<TF>(c, f, sub_dt));
}

int mhh_pres_check_divergence(mhh_ctx* ctx, int swpres, const mhh_fields* f, double* divmax)
{
    if (!ctx || !f || !divmax) return MHH_E_INVALID;
    if (swpres != 2 && swpres != 4) { ctx->err = "pres_check_divergence: swpres must be 2 or 4"; return MHH_E_INVALID; }
    if (swpres == 4)
    {
        if (ctx->dtype == MHH_F64) return pres4_div_impl<double>(static_cast<Ctx<double>*>(ctx), f, divmax);
        return pres4_div_impl<float>(static_cast<Ctx<float>*>(ctx), f, divmax);
    }
    if (ctx->dtype == MHH_F64) { typedef double TF; Ctx<TF>* c = static_cast<Ctx<TF>*>(ctx); NEED_BASE(c); return reduce_impl<TF, 2>(c, P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), 0, 0, 0, divmax); }
    else { typedef float TF; Ctx<TF>* c = static_cast<Ctx<TF>*>(ctx); NEED_BASE(c); return reduce_impl<TF, 2>(c, P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), 0, 0, 0, divmax); }
}

int mhh_pres_fft_roundtrip(mhh_ctx* ctx, const void* in_compact, void* out_compact, int solve)
{ DISPATCH1(ctx, fft_roundtrip_impl<TF>(c, P<TF>(in_compact), P<TF>(out_compact), solve)); }

int mhh_timeloop_rk3(mhh_ctx* ctx, void* a, void* at, int substep, double dt)
{ DISPATCH1(ctx, rk3_impl<TF>(c, P<TF>(a), P<TF>(at), substep, dt)); }

int mhh_dycore_substep(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, int substep, double dt)
{
    if (!f) return MHH_E_INVALID;
    DISPATCH1(ctx, substep_impl<TF>(c, f, prm, substep, dt));
}

int mhh_dycore_substep_pre(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm)
{
    if (!f) return MHH_E_INVALID;
    DISPATCH1(ctx, substep_pre_impl<TF>(c, f, prm));
}

int mhh_dycore_set_ghost_cells(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm)
{
    if (!f || !prm) return MHH_E_INVALID;
    DISPATCH1(ctx, ghost_all_impl<TF>(c, f, prm));
}

int mhh_dycore_tendencies(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm)
{
    if (!f) return MHH_E_INVALID;
    DISPATCH1(ctx, tendencies_impl<TF>(c, f, prm));
}

int mhh_dycore_substep_post(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, int substep, double dt)
{
    if (!f) return MHH_E_INVALID;
    DISPATCH1(ctx, substep_post_impl<TF>(c, f, prm, substep, dt));
}

int mhh_dycore_step(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, double dt)
{
    for (int ss = 0; ss < 3; ++ss)
    {
        int rc = mhh_dycore_substep(ctx, f, prm, ss, dt);
        if (rc != MHH_OK) return rc;
    }
    return MHH_OK;
}

int mhh_dycore_step_host(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, double dt, int nsteps,
                         void* h_u, void* h_v, void* h_w, void* const* h_s)
{
    if (!f || !prm) return MHH_E_INVALID;
    DISPATCH1(ctx, step_host_impl<TF>(c, f, prm, dt, nsteps, h_u, h_v, h_w, h_s));
}

} // extern "C"
