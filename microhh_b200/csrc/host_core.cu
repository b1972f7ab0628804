// mhhb200 -- C ABI implementation (see include/mhhb200.h).  Host-side orchestration only:
// argument checks, launch configuration, the per-context tables.  No CPU compute path exists.
#include "host_common.cuh"

NcclApi* nccl_api(std::string& err)
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

void prof_mark(mhh_ctx* c, const char* name)
{
    if (!c->prof) return;
    cudaEvent_t e;
    if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
    else if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, c->stream);
    c->prof_events.emplace_back(name, e);
}

namespace mhhhost {

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
int create_impl(const mhh_grid_desc* d, int dtype, int device, mhh_ctx** out)
{
    Ctx<TF>* c = new (std::nothrow) Ctx<TF>();
    if (!c) return MHH_E_NOMEM;
    *out = c;
    c->dtype = dtype; c->device = device; c->desc = *d;
    { const char* e = getenv("MHH_FORCE_PLAIN"); c->force_plain = e && e[0] == '1'; }
    { const char* e = getenv("MHH_NO_TMA"); c->no_tma = e && e[0] == '1'; }
    { const char* e = getenv("MHH_FUSE_SCALAR"); if (e) c->fuse_scalar = e[0] == '1'; }
    { const char* e = getenv("MHH_FUSE_ADVEC2"); if (e) c->fuse_advec2 = e[0] == '1'; }
    { const char* e = getenv("MHH_SCAL_TMA"); if (e) c->scal_tma = e[0] == '1'; }
    { const char* e = getenv("MHH_GRAPH"); if (e) { c->use_graph = e[0] == '1'; c->graph_mode = c->use_graph ? 1 : 0; } }
    { const char* e = getenv("MHH_TILE3_Y"); if (e) { int v = atoi(e); if (v == 3 || v == 4 || v == 5) c->tile3_y = v; } }
    { const char* e = getenv("MHH_EVISC_MB"); if (e) c->evisc_mb = atoi(e); }
    { const char* e = getenv("MHH_EVISC_TMA"); if (e) c->evisc_tma = atoi(e) != 0; }
    { const char* e = getenv("MHH_STREAM_VEC2"); if (e) c->stream_vec2 = atoi(e) != 0; }
    { const char* e = getenv("MHH_EVISC3_MB"); if (e) c->evisc3_mb = atoi(e); }
    { const char* e = getenv("MHH_EVISC3_NPL"); if (e) c->evisc3_npl = atoi(e); }
    { const char* e = getenv("MHH_PREFETCH"); if (e) c->prefetch = std::max(0, std::min(8, atoi(e))); }
    { const char* e = getenv("MHH_TILE_Y"); if (e && atoi(e) == 16) c->tile_y = 16; else if (e && atoi(e) == 8) c->tile_y = 8; }
    CUDA_TRY(c, cudaSetDevice(device));
    int nsm = 0;
    CUDA_TRY(c, cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    c->num_sms = nsm;
    CUDA_TRY(c, cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    // (default priority: with the highest one the update runs first and the x transform simply takes that much longer --
    // measured fp64 44.8 either way, fp32 28.65 vs 27.96 with the default priority, 28.8 without the fork)
    CUDA_TRY(c, cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
    CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    { const char* e = getenv("MHH_OVERLAP"); if (e) c->overlap = atoi(e) != 0; }

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
    c->h_dzhi.assign(static_cast<const TF*>(d->dzhi), static_cast<const TF*>(d->dzhi) + kc);
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

    CUDA_TRY(c, cudaMalloc(&c->d_mlen0, sizeof(TF) * kc));
    int rcp = pres_create<TF>(c);
    // the metric arrays were copied: do not keep the caller's host pointers
    c->desc.z = c->desc.zh = c->desc.dz = c->desc.dzh = c->desc.dzi = c->desc.dzhi = nullptr; c->desc.dzi4 = c->desc.dzhi4 = nullptr;
    return rcp;
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
    if (thref) c->h_thref.assign(static_cast<const TF*>(thref), static_cast<const TF*>(thref) + kc);
    if (threfh) c->h_threfh.assign(static_cast<const TF*>(threfh), static_cast<const TF*>(threfh) + kc);

    {
        std::vector<TF> mlen0(kc, TF(0));
        const TF* dz = c->h_dz.data();
        // Smagorinsky filter width per level: mlen0 = (dx*dy*dz)^(1/3) (cs applied at call time)
        for (int k = 0; k < kc; ++k)
            mlen0[k] = std::pow(g.dx * g.dy * dz[k], TF(1. / 3.));
        CUDA_TRY(c, cudaMemcpy(c->d_mlen0, mlen0.data(), sizeof(TF) * kc, cudaMemcpyHostToDevice));
    }
    int rcp = pres_set_values<TF>(c);
    if (rcp != MHH_OK) return rcp;
    c->basestate_set = true;
    return MHH_OK;
}


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

// two cells per thread in the streaming kernels: even icells (every row starts on a vector boundary) and vector-aligned arrays
template <typename TF>
static bool stream_vec2(const Ctx<TF>* c, std::initializer_list<const void*> ptrs)
{
    if (!c->stream_vec2 || (c->g.icells & 1)) return false;
    for (const void* p : ptrs)
        if (reinterpret_cast<uintptr_t>(p) % (2 * sizeof(TF)) != 0) return false;
    return true;
}

template <typename TF>
int rk3_impl(Ctx<TF>* c, TF* a, TF* at, int substep, double dt, cudaStream_t st)
{
    const GridDev<TF>& g = c->g;
    if (!st) st = c->stream;
    NEED(c, a, "field"); NEED(c, at, "tendency");
    if (substep < 0 || substep > 2) { c->err = "substep must be 0..2"; return MHH_E_INVALID; }
    const TF cA[3] = {TF(0.), TF(-5. / 9.), TF(-153. / 128.)};
    const TF cB[3] = {TF(1. / 3.), TF(15. / 16.), TF(8. / 15.)};
    const int nxt = (substep + 1) % 3;
    if (stream_vec2<TF>(c, {a, at}))
        rk3_v2_kernel<TF><<<dim3((g.icells / 2 + 63) / 64, (g.jcells + 3) / 4, g.kcells), c->blk(), 0, st>>>(a, at, cB[substep] * (TF)dt, cA[nxt], nxt == 0, g);
    else
        rk3_kernel<TF><<<c->grd_all(), c->blk(), 0, st>>>(a, at, cB[substep] * (TF)dt, cA[nxt], nxt == 0, g);
    // on the side stream the kernel runs under the pressure solve: its own time does not show between two marks of the main stream
    if (st == c->stream) KCHECKN(c, "rk3_kernel"); else KCHECKN(c, "rk3_kernel_overlapped");
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
    // thermo on a 4th-order grid is Thermo_buoy (every shipped swspatialorder = 4 case with thermo: drycbl, drycblslope,
    // prandtlslope, rayleighbenard, rayleightaylor, vanheerwaarden2016, weakscaling); Thermo_dry's kernels are 2nd-order only
    if (prm->swthermo != 0 && prm->swthermo != 2) { c->err = "the 4th-order sub-step couples to swthermo = buoy (2) only"; return MHH_E_INVALID; }
    if (prm->swthermo == 2 && !c->buoy_set) { c->err = "dycore_substep: swthermo = buoy needs mhh_dycore_set_thermo_buoy"; return MHH_E_INVALID; }
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
    if (prm->swthermo == 2)
    {
        // thermo.exec (src/model.cxx:388) still sees the normal-type w ghost cells of Boundary::set_ghost_cells
        if ((rc = ghost4w_impl<TF>(c, P<TF>(f->w), 0)) != MHH_OK) return rc;
        if ((rc = thermo_buoy_impl<TF>(c, f, &c->buoy)) != MHH_OK) return rc;
    }
    if ((rc = ghost4w_impl<TF>(c, P<TF>(f->w), 1)) != MHH_OK) return rc;        // (without thermo the normal-type fill before it would only be overwritten)
    if ((rc = o4_impl<TF>(c, f, prm->swadvec, false)) != MHH_OK) return rc;
    if ((rc = ghost4w_impl<TF>(c, P<TF>(f->w), 0)) != MHH_OK) return rc;
    if ((rc = o4_impl<TF>(c, f, 0, true)) != MHH_OK) return rc;
    const double cBd[3] = {1. / 3., 15. / 16., 8. / 15.};
    if (c->forcing_set)
    {
        if ((rc = buffer_exec_impl<TF>(c, f, &c->forcing)) != MHH_OK) return rc;
        if ((rc = force_exec_impl<TF>(c, f, &c->forcing, cBd[substep] * dt)) != MHH_OK) return rc;
    }
    if ((rc = ghost4w_impl<TF>(c, P<TF>(f->w), 1)) != MHH_OK) return rc;
    if ((rc = pres4_exec_impl<TF>(c, f, cBd[substep] * dt)) != MHH_OK) return rc;
    if ((rc = ghost4w_impl<TF>(c, P<TF>(f->w), 0)) != MHH_OK) return rc;
    TF* tend[3 + MHH_MAX_SCALARS] = {P<TF>(f->ut), P<TF>(f->vt), P<TF>(f->wt)};
    for (int n = 0; n < f->ns; ++n) tend[3 + n] = P<TF>(f->st[n]);
    for (int n = 0; n < 3 + f->ns; ++n)
        if ((rc = rk3_impl<TF>(c, prog[n], tend[n], substep, dt, nullptr)) != MHH_OK) return rc;
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
    if (prm->swadvec == 4 || prm->swadvec == 41 || prm->swdiff == 4)
    {
        if ((prm->swadvec != 4 && prm->swadvec != 41) || prm->swdiff != 4)
        { c->err = "dycore_substep: the 4th-order configuration is swadvec = 4 or 4m (41) with swdiff = 4 (and pres_4)"; return MHH_E_INVALID; }
        o4 = true;
        return MHH_OK;
    }
    if ((prm->swadvec != 25 && prm->swadvec != 2 && prm->swadvec != 24 && prm->swadvec != 262) || (prm->swdiff != 1 && prm->swdiff != 2 && prm->swdiff != 3))
    { c->err = "dycore_substep: swadvec must be 2i5 (25), 2, 2i4 (24), 2i62 (262), 4 or 4m (41), swdiff smag2 (1), 2, tke2 (3) or 4"; return MHH_E_INVALID; }
    const bool smag = prm->swdiff == 1 || prm->swdiff == 3;
    if (prm->swthermo < 0 || prm->swthermo > 3) { c->err = "dycore_substep: swthermo must be 0, dry (1), buoy (2) or moist (3)"; return MHH_E_INVALID; }
    if (prm->swthermo == 3)
    {
        if (!c->moist_set) { c->err = "dycore_substep: swthermo = moist needs mhh_dycore_set_thermo_moist"; return MHH_E_INVALID; }
        // the closures derive N2 from scalar 0 and the context's thref (= thvref for Thermo_moist, src/thermo_moist.cxx:459-475)
        if (smag && c->moist.ithl != 0) { c->err = "dycore_substep: swthermo = moist with an LES closure needs thl as scalar 0"; return MHH_E_INVALID; }
        if (prm->swdiff == 3) { c->err = "dycore_substep: swthermo = moist goes with swdiff = smag2 or 2 (Diff_tke2 is wired to Thermo_dry)"; return MHH_E_INVALID; }
    }
    if (prm->swthermo == 2)
    {
        // the eddy-viscosity kernels derive N2 and the surface buoyancy gradient the Thermo_dry way; no shipped case pairs buoy with an LES closure
        if (smag) { c->err = "dycore_substep: swthermo = buoy goes with swdiff = 2 or 4 (the LES closures are wired to Thermo_dry)"; return MHH_E_INVALID; }
        if (!c->buoy_set) { c->err = "dycore_substep: swthermo = buoy needs mhh_dycore_set_thermo_buoy"; return MHH_E_INVALID; }
    }
    int rc = check_mom<TF>(c, f, smag, smag && prm->surface_model != 0);
    if (rc != MHH_OK) return rc;
    if (prm->swdiff == 3)
    {
        if (!c->tke2_set) { c->err = "dycore_substep: swdiff = tke2 needs mhh_dycore_set_tke2"; return MHH_E_INVALID; }
        if ((rc = tke2_check<TF>(c, f, prm, &c->tke2)) != MHH_OK) return rc;
    }
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
    if (prm->swdiff == 3 && (rc = tke2_visc_impl<TF>(c, f, prm, &c->tke2, nullptr)) != MHH_OK) return rc;
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
    const bool smag = prm->swdiff == 1 || prm->swdiff == 3, adv5 = prm->swadvec == 25, buoy = prm->swthermo == 1;
    const mhh_tke2* tke = prm->swdiff == 3 ? &c->tke2 : nullptr;                             // Diff_tke2::exec = the smag2 kernels, evisc per scalar
    if (prm->swthermo == 2 && (rc = thermo_buoy_impl<TF>(c, f, &c->buoy)) != MHH_OK) return rc;  // Thermo_buoy::exec, ahead of advec.exec like Model::exec
    if (prm->swthermo == 3 && (rc = thermo_moist_impl<TF>(c, f, &c->moist)) != MHH_OK) return rc; // Thermo_moist::exec (base-state update + buoyancy)
    if (prm->swadvec == 24 || prm->swadvec == 262)                                           // 2i4 | 2i62: the advection alone, then the diffusion (+ buoyancy)
    {
        // Model::exec order: thermo.exec (buoyancy alone), advec.exec, diff.exec -- the diffusion-only launches carry no buoyancy
        if (buoy && (rc = o2_impl<TF>(c, f, false, false, true)) != MHH_OK) return rc;
        if ((rc = adv2i_impl<TF>(c, f, prm->swadvec)) != MHH_OK) return rc;
        rc = smag ? tend_impl<TF>(c, f, prm, false, true, false, tke) : o2_impl<TF>(c, f, false, true, false);
    }
    else if (adv5 && smag) rc = tend_impl<TF>(c, f, prm, true, true, buoy, tke);             // 2i5 + smag2 | tke2 (+ buoyancy)
    else if (!adv5 && !smag) rc = o2_impl<TF>(c, f, true, true, buoy);                       // 2 + 2 (+ buoyancy)
    else if (!adv5)                                                                          // 2 + smag2 (drycblles as shipped)
    {
        // one fused TMA-staged kernel when it applies (igc >= 3, at most the one scalar that rides along) ...
        rc = c->fuse_advec2 ? tend_impl<TF>(c, f, prm, true, true, buoy, tke, true) : MHH_NOT_FUSED;
        if (rc == MHH_NOT_FUSED)
        {
            // ... else Advec_2 point-wise, then the diffusion alone
            if ((rc = o2_impl<TF>(c, f, true, false, buoy)) != MHH_OK) return rc;
            rc = tend_impl<TF>(c, f, prm, false, true, false, tke);
        }
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
    // buffer.exec and force.exec ("keep this one always right before the pressure", src/model.cxx:416-430) when registered
    if (c->forcing_set)
    {
        const double cBf[3] = {1. / 3., 15. / 16., 8. / 15.};
        if ((rc = buffer_exec_impl<TF>(c, f, &c->forcing)) != MHH_OK) return rc;
        if ((rc = force_exec_impl<TF>(c, f, &c->forcing, cBf[substep] * dt)) != MHH_OK) return rc;
    }
    // limiter.exec on sgstke: the reference applies it "as the last tendency", after pres.exec (src/model.cxx:439-440); the
    // pressure solve neither reads nor writes a scalar tendency, so it runs here, ahead of the forked scalar update
    if (prm->swdiff == 3)
    {
        const double cBl[3] = {1. / 3., 15. / 16., 8. / 15.};
        const int n = c->tke2.isgstke;
        if ((rc = limiter_impl<TF>(c, P<TF>(f->st[n]), P<TF>(f->s[n]), TF(MHH_SGSTKE_MIN), (TF)(cBl[substep] * dt))) != MHH_OK) return rc;
    }
    // 4. pres.exec (solve), then pressure correction fused with timeloop.exec
    const TF cA[3] = {TF(0.), TF(-5. / 9.), TF(-153. / 128.)};
    const TF cB[3] = {TF(1. / 3.), TF(15. / 16.), TF(8. / 15.)};
    const double cBd[3] = {1. / 3., 15. / 16., 8. / 15.};
    const double sub_dt = cBd[substep] * dt;          // Timeloop::get_sub_time_step (double)
    // The scalars are complete once their tendencies are: nothing in pres.exec reads or writes them.  Their RK3 update is forked
    // onto the side stream here and runs under the (latency-bound) Poisson kernels; the main stream joins at the end of the
    // sub-step, before anything reads the scalars again.
    const bool fork = c->overlap && c->side_stream && f->ns > 0;
    if (fork)
    {
        CUDA_TRY(c, cudaEventRecord(c->ev_fork, c->stream));
        CUDA_TRY(c, cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
        for (int n = 0; n < f->ns; ++n)
            if ((rc = rk3_impl<TF>(c, P<TF>(f->s[n]), P<TF>(f->st[n]), substep, dt, c->side_stream)) != MHH_OK) return rc;
        CUDA_TRY(c, cudaEventRecord(c->ev_join, c->side_stream));
    }
    if ((rc = pres_solve_impl<TF>(c, f, sub_dt)) != MHH_OK) return rc;
    const int nxt = (substep + 1) % 3;
    PresArgs<TF> a{P<TF>(f->ut), P<TF>(f->vt), P<TF>(f->wt), P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), P<TF>(f->p)};
    if (stream_vec2<TF>(c, {a.ut, a.vt, a.wt, a.u, a.v, a.w, a.p}))
        pres_out_rk3_v2_kernel<TF><<<dim3((g.icells / 2 + 63) / 64, (g.jcells + 3) / 4, g.kcells), c->blk(), 0, c->stream>>>(a, P<TF>(f->u), P<TF>(f->v), P<TF>(f->w),
                cB[substep] * (TF)dt, cA[nxt], nxt == 0, g);
    else
        pres_out_rk3_kernel<TF><<<c->grd_all(), c->blk(), 0, c->stream>>>(a, P<TF>(f->u), P<TF>(f->v), P<TF>(f->w),
                cB[substep] * (TF)dt, cA[nxt], nxt == 0, g);
    KCHECKN(c, "pres_out_rk3_kernel");
    if (fork) CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    else
        for (int n = 0; n < f->ns; ++n)
            if ((rc = rk3_impl<TF>(c, P<TF>(f->s[n]), P<TF>(f->st[n]), substep, dt, nullptr)) != MHH_OK) return rc;
    return MHH_OK;
}

// ---- Buffer (damping layer) and Force (large-scale forcings) ----------------------------------------------------------
// Buffer<TF>::exec (src/buffer.cxx:170-205): u, v and the scalars from the first full level inside the layer, w from the first
// half level; the damping factor sigma ((z - zstart)/(zsize - zstart))^beta is tabulated per level.
template <typename TF>
int buffer_exec_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_forcing* fo)
{
    const GridDev<TF>& g = c->g;
    NEED(c, fo, "forcing"); NEED(c, f, "fields");
    if (!fo->swbuffer) return MHH_OK;
    const int kc = g.kcells;
    if (!c->d_sigmaz) CUDA_TRY(c, cudaMalloc(&c->d_sigmaz, sizeof(TF) * 2 * kc));
    if (c->buf_key[0] != fo->buffer_zstart || c->buf_key[1] != fo->buffer_sigma || c->buf_key[2] != fo->buffer_beta)
    {
        // Buffer::create (src/buffer.cxx:103-124) and the per-level factors of calc_buffer (:46-49)
        const TF zstart = (TF)fo->buffer_zstart, sigma = (TF)fo->buffer_sigma, beta = (TF)fo->buffer_beta;
        const TF zsizebuf = g.zsize - zstart;
        std::vector<TF> zh(kc), sg(2 * (size_t)kc, TF(0));
        CUDA_TRY(c, cudaMemcpy(zh.data(), g.zh, sizeof(TF) * kc, cudaMemcpyDeviceToHost));
        int kb = g.kstart, kbh = g.kstart;
        for (int k = g.kstart; k < g.kend; ++k) { if (c->h_z[k] < zstart) ++kb; if (zh[k] < zstart) ++kbh; }
        if (kbh == g.kend) { c->err = "Buffer is too close to the model top"; return MHH_E_INVALID; }
        for (int k = kb; k < g.kend; ++k) sg[k] = sigma * std::pow((c->h_z[k] - zstart) / zsizebuf, beta);
        for (int k = kbh; k < g.kend; ++k) sg[kc + k] = sigma * std::pow((zh[k] - zstart) / zsizebuf, beta);
        CUDA_TRY(c, cudaMemcpyAsync(c->d_sigmaz, sg.data(), sizeof(TF) * 2 * kc, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        c->buf_k = kb; c->buf_kh = kbh;
        c->buf_key[0] = fo->buffer_zstart; c->buf_key[1] = fo->buffer_sigma; c->buf_key[2] = fo->buffer_beta;
    }
    dim3 b = c->blk();
    auto launch = [&](TF* at, const TF* a, const TF* prof, bool half) -> int {
        if (!prof) return MHH_OK;                          // no profile: the field is not damped
        NEED(c, at, "tendency"); NEED(c, a, "field");
        const int k0 = half ? c->buf_kh : c->buf_k;
        if (k0 >= g.kend) return MHH_OK;
        dim3 gr((g.imax + 63) / 64, (g.jmax + 3) / 4, g.kend - k0);
        buffer_kernel<TF><<<gr, b, 0, c->stream>>>(at, a, prof, c->d_sigmaz + (half ? kc : 0), k0, g);
        KCHECKN(c, "buffer_kernel");
        return MHH_OK; };
    int rc;
    if ((rc = launch(P<TF>(f->ut), P<TF>(f->u), P<TF>(fo->bufferprof_u), false)) != MHH_OK) return rc;
    if ((rc = launch(P<TF>(f->vt), P<TF>(f->v), P<TF>(fo->bufferprof_v), false)) != MHH_OK) return rc;
    if ((rc = launch(P<TF>(f->wt), P<TF>(f->w), P<TF>(fo->bufferprof_w), true)) != MHH_OK) return rc;
    for (int n = 0; n < f->ns; ++n)
        if ((rc = launch(P<TF>(f->st[n]), P<TF>(f->s[n]), P<TF>(fo->bufferprof_s[n]), false)) != MHH_OK) return rc;
    return MHH_OK;
}

// Force<TF>::exec (src/force.cxx:608-700): large-scale pressure force (fixed mass flux / pressure gradient / geostrophic wind
// with Coriolis), large-scale sources and local subsidence of the scalars
template <typename TF>
int force_exec_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_forcing* fo, double sub_dt)
{
    const GridDev<TF>& g = c->g;
    NEED(c, fo, "forcing"); NEED(c, f, "fields");
    dim3 b = c->blk(), gr = c->grd_interior();
    if (fo->swlspres == MHH_LSPRES_UFLUX)
    {
        NEED(c, f->u, "u"); NEED(c, f->ut, "ut");
        if (!c->d_sums) CUDA_TRY(c, cudaMalloc(&c->d_sums, 2 * sizeof(double)));
        CUDA_TRY(c, cudaMemsetAsync(c->d_sums, 0, 2 * sizeof(double), c->stream));
        mean_uut_kernel<TF><<<gr, b, 0, c->stream>>>(P<TF>(f->u), P<TF>(f->ut), g, c->d_sums);
        KCHECKN(c, "mean_uut_kernel");
        if (c->nranks > 1)
        {
            if (!c->comm) { c->err = "slab context without communicator: call mhh_comm_init first"; return MHH_E_INVALID; }
            NcclApi* api = nccl_api(c->err);
            if (!api) return MHH_E_CUDA;
            NCCL_TRY(c, api, api->AllReduce(c->d_sums, c->d_sums, 2, ncclFloat64, ncclSum, c->comm, c->stream));
        }
        const double inv_vol = 1. / ((double)g.itot * (double)g.jtot * (double)g.zsize);
        body_force_kernel<TF, true><<<gr, b, 0, c->stream>>>(P<TF>(f->ut), c->d_sums, inv_vol, (TF)fo->uflux, (TF)fo->utrans, (TF)sub_dt, TF(0), g);
        KCHECKN(c, "body_force_kernel");
    }
    else if (fo->swlspres == MHH_LSPRES_DPDX)
    {
        NEED(c, f->ut, "ut");
        body_force_kernel<TF, false><<<gr, b, 0, c->stream>>>(P<TF>(f->ut), nullptr, 0., TF(0), TF(0), TF(1), (TF)(-1. * fo->dpdx), g);
        KCHECKN(c, "body_force_kernel");
    }
    else if (fo->swlspres == MHH_LSPRES_GEO)
    {
        NEED(c, f->u, "u"); NEED(c, f->v, "v"); NEED(c, f->ut, "ut"); NEED(c, f->vt, "vt"); NEED(c, fo->ug, "ug"); NEED(c, fo->vg, "vg");
        if (g.dzi4) coriolis_kernel<TF, 4><<<gr, b, 0, c->stream>>>(P<TF>(f->ut), P<TF>(f->vt), P<TF>(f->u), P<TF>(f->v), P<TF>(fo->ug), P<TF>(fo->vg), (TF)fo->fc, (TF)fo->utrans, (TF)fo->vtrans, g);
        else coriolis_kernel<TF, 2><<<gr, b, 0, c->stream>>>(P<TF>(f->ut), P<TF>(f->vt), P<TF>(f->u), P<TF>(f->v), P<TF>(fo->ug), P<TF>(fo->vg), (TF)fo->fc, (TF)fo->utrans, (TF)fo->vtrans, g);
        KCHECKN(c, "coriolis_kernel");
    }
    else if (fo->swlspres != MHH_LSPRES_OFF) { c->err = "force: unknown swlspres"; return MHH_E_INVALID; }
    for (int n = 0; n < f->ns; ++n)
    {
        const TF* sls = P<TF>(fo->ls_s[n]);
        const TF* wls = P<TF>(fo->wls);
        if (!sls && !wls) continue;
        NEED(c, f->s[n], "scalar"); NEED(c, f->st[n], "scalar tendency");
        scalar_forcing_kernel<TF><<<gr, b, 0, c->stream>>>(P<TF>(f->st[n]), P<TF>(f->s[n]), sls, wls, g);
        KCHECKN(c, "scalar_forcing_kernel");
    }
    return MHH_OK;
}

// ---- Boundary_surface (Monin-Obukhov surface model, constant z0, lookup solver) --------------------------------------
// bsk::prepare_lut (include/boundary_surface_kernels.h:78-138), in the reference's mix of TF / double / float arithmetic
template <typename TF>
int surface_init_impl(Ctx<TF>* c, double z0m_d, double z0h_d, int mbcbot, int thermobc)
{
    if (mbcbot != 0) { c->err = "boundary_surface: only mbcbot = noslip (Dirichlet) is implemented"; return MHH_E_INVALID; }
    if (thermobc != 0 && thermobc != 2) { c->err = "boundary_surface: thermobc must be Dirichlet (0) or flux (2)"; return MHH_E_INVALID; }
    const int nlut = SURF_NLUT;
    const TF z0m = (TF)z0m_d, z0h = (TF)z0h_d, zsl = c->h_z[c->g.kstart];
    std::vector<TF> zL_tmp(nlut);
    const TF zL_max = 10., zL_min = -1.e4, zLrange_min = -5.;
    TF dzL = (zL_max - zLrange_min) / (9. * nlut / 10. - 1.);
    zL_tmp[0] = -zL_max;
    for (int n = 1; n < 9 * nlut / 10; ++n) zL_tmp[n] = zL_tmp[n - 1] + dzL;
    const TF zLend = -(zL_min - zLrange_min);
    TF r = 1.01, r0 = 1.e30;
    while (std::abs((r - r0) / r0) > 1.e-10)
    {
        r0 = r;
        r = std::pow(1. - (zLend / dzL) * (1. - r), (1. / (nlut / 10.)));
    }
    for (int n = 9 * nlut / 10; n < nlut; ++n) { zL_tmp[n] = zL_tmp[n - 1] + dzL; dzL *= r; }
    std::vector<float> zL(nlut), fs(nlut);
    for (int n = 0; n < nlut; ++n) zL[n] = -zL_tmp[nlut - n - 1];
    // host twins of the stability functions (include/monin_obukhov.h)
    auto psim = [](TF zeta, bool unstable) -> TF {
        if (unstable) { const TF ph = std::pow(TF(1.) + TF(3.6) * std::pow(std::abs(zeta), TF(2. / 3.)), TF(-1. / 2.)); return TF(3.) * std::log((TF(1.) + TF(1.) / ph) / TF(2.)); }
        const TF a = 1, b = TF(2) / TF(3), cc = 5, dd = TF(0.35);
        return -b * (zeta - (cc / dd)) * std::exp(-dd * zeta) - a * zeta - (b * cc) / dd; };
    auto psih = [](TF zeta, bool unstable) -> TF {
        if (unstable) { const TF ph = std::pow(TF(1.) + TF(7.9) * std::pow(std::abs(zeta), TF(2. / 3.)), TF(-1. / 2.)); return TF(3.) * std::log((TF(1.) + TF(1.) / ph) / TF(2.)); }
        const TF a = 1, b = TF(2) / TF(3), cc = 5, dd = TF(0.35);
        return -b * (zeta - (cc / dd)) * std::exp(-dd * zeta) - std::pow(TF(1) + b * a * zeta, TF(1.5)) - (b * cc) / dd + TF(1); };
    auto fm = [&](TF L) -> TF { const bool u = L <= TF(0.); return TF(0.4) / (std::log(zsl / z0m) - psim(zsl / L, u) + psim(z0m / L, u)); };
    auto fh = [&](TF L) -> TF { const bool u = L <= TF(0.); return TF(0.4) / (std::log(zsl / z0h) - psih(zsl / L, u) + psih(z0h / L, u)); };
    for (int n = 0; n < nlut; ++n)
    {
        const TF L = zsl / zL[n];
        if (thermobc == 2) fs[n] = zL[n] * std::pow(fm(L), 3);
        else fs[n] = zL[n] * std::pow(fm(L), 2) / fh(L);
    }
    if (!c->d_zL_sl) { CUDA_TRY(c, cudaMalloc(&c->d_zL_sl, sizeof(float) * nlut)); CUDA_TRY(c, cudaMalloc(&c->d_f_sl, sizeof(float) * nlut)); }
    CUDA_TRY(c, cudaMemcpy(c->d_zL_sl, zL.data(), sizeof(float) * nlut, cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_f_sl, fs.data(), sizeof(float) * nlut, cudaMemcpyHostToDevice));
    c->surf_mbcbot = mbcbot; c->surf_thermobc = thermobc;
    return MHH_OK;
}

// Boundary_surface<TF>::exec (src/boundary_surface.cxx:836-990), Thermo_dry or Thermo_disabled
template <typename TF>
int surface_exec_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, const mhh_surface* s)
{
    NEED_BASE(c);
    const GridDev<TF>& g = c->g;
    NEED(c, s, "surface"); NEED(c, f, "fields");
    if (!c->d_zL_sl) { c->err = "mhh_boundary_surface_init has not been called"; return MHH_E_INVALID; }
    if (g.igc < 2 || g.jgc < 2) { c->err = "boundary_surface: the wind filter needs igc, jgc >= 2"; return MHH_E_INVALID; }
    if (prm->swthermo > 1) { c->err = "boundary_surface on the device: Thermo_dry or no thermo (swthermo = buoy / moist: keep the host's surface model, mhh_dycore_substep_pre / _post)"; return MHH_E_INVALID; }
    const bool neutral = prm->swthermo == 0;
    SurfArgs<TF> a{};
    a.ustar = P<TF>(s->ustar); a.obuk = P<TF>(s->obuk); a.nobuk = static_cast<int*>(s->nobuk); a.dutot = P<TF>(s->dutot);
    a.z0m = P<TF>(s->z0m); a.z0h = P<TF>(s->z0h); a.zL_sl = c->d_zL_sl; a.f_sl = c->d_f_sl;
    NEED(c, a.ustar, "ustar"); NEED(c, a.obuk, "obuk"); NEED(c, a.nobuk, "nobuk"); NEED(c, a.dutot, "dutot"); NEED(c, a.z0m, "z0m"); NEED(c, a.z0h, "z0h");
    a.u = P<TF>(f->u); a.v = P<TF>(f->v); a.ubot = P<TF>(f->u_bot); a.vbot = P<TF>(f->v_bot);
    a.ufluxbot = P<TF>(f->u_fluxbot); a.vfluxbot = P<TF>(f->v_fluxbot); a.ugradbot = P<TF>(f->u_gradbot); a.vgradbot = P<TF>(f->v_gradbot);
    a.dudz = P<TF>(f->dudz_mo); a.dvdz = P<TF>(f->dvdz_mo); a.dbdz = P<TF>(f->dbdz_mo);
    NEED(c, a.u, "u"); NEED(c, a.v, "v"); NEED(c, a.ubot, "u_bot"); NEED(c, a.vbot, "v_bot"); NEED(c, a.ufluxbot, "u_fluxbot"); NEED(c, a.vfluxbot, "v_fluxbot");
    NEED(c, a.ugradbot, "u_gradbot"); NEED(c, a.vgradbot, "v_gradbot"); NEED(c, a.dudz, "dudz_mo"); NEED(c, a.dvdz, "dvdz_mo");
    if (f->ns < 0 || f->ns > MHH_MAX_SCALARS) { c->err = "ns out of range"; return MHH_E_INVALID; }
    a.ns = f->ns;
    for (int n = 0; n < f->ns; ++n)
    {
        a.s[n] = P<TF>(f->s[n]); a.sbot[n] = P<TF>(f->s_bot[n]); a.sgradbot[n] = P<TF>(f->s_gradbot[n]); a.sfluxbot[n] = P<TF>(f->s_fluxbot[n]);
        a.sbc[n] = s->sbcbot[n];
        NEED(c, a.s[n], "scalar"); NEED(c, a.sbot[n], "s_bot"); NEED(c, a.sgradbot[n], "s_gradbot"); NEED(c, a.sfluxbot[n], "s_fluxbot");
    }
    a.mbcbot = c->surf_mbcbot; a.thermobc = c->surf_thermobc; a.neutral = neutral ? 1 : 0;
    if (!neutral)
    {
        if (f->ns < 1) { c->err = "boundary_surface: thermo dry needs scalar 0 (th)"; return MHH_E_INVALID; }
        if (c->h_thref.empty() || c->h_threfh.empty()) { c->err = "boundary_surface: thref / threfh were not given to mhh_set_basestate"; return MHH_E_INVALID; }
        NEED(c, a.dbdz, "dbdz_mo");
        if (s->sbcbot[0] != c->surf_thermobc) { c->err = "boundary_surface: sbcbot[0] differs from the thermobc the lookup table was built for"; return MHH_E_INVALID; }
        a.thref = c->h_thref[g.kstart]; a.threfh = c->h_threfh[g.kstart];
        a.gthref = TF(GRAV) / a.thref; a.gthrefh = TF(GRAV) / a.threfh;
    }
    a.zsl = c->h_z[g.kstart];
    dim3 b(64, 4), gi((g.imax + 63) / 64, (g.jmax + 3) / 4), ga((g.icells + 63) / 64, (g.jcells + 3) / 4);
    int rc;
    surface_dutot_kernel<TF><<<gi, b, 0, c->stream>>>(a, g);
    KCHECKN(c, "surface_dutot_kernel");
    if ((rc = cyclic_impl<TF>(c, a.dutot, MHH_EDGE_BOTH, true)) != MHH_OK) return rc;
    surface_stability_kernel<TF><<<ga, b, 0, c->stream>>>(a, g);
    KCHECKN(c, "surface_stability_kernel");
    surface_flux_kernel<TF><<<gi, b, 0, c->stream>>>(a, g);
    KCHECKN(c, "surface_flux_kernel");
    if ((rc = cyclic_impl<TF>(c, a.ufluxbot, MHH_EDGE_BOTH, true)) != MHH_OK) return rc;
    if ((rc = cyclic_impl<TF>(c, a.vfluxbot, MHH_EDGE_BOTH, true)) != MHH_OK) return rc;
    surface_values_kernel<TF><<<ga, b, 0, c->stream>>>(a, g);
    KCHECKN(c, "surface_values_kernel");
    return MHH_OK;
}

// The self-driven LES sub-step in the reference's order (src/model.cxx:368-504): cyclic + ghost cells -> exec_viscosity ->
// boundary.exec (surface model) -> set_ghost_cells -> thermo + advec + diff -> pres -> timeloop
template <typename TF>
int substep_surface_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, const mhh_surface* s, int substep, double dt)
{
    int rc;
    if ((rc = substep_pre_impl<TF>(c, f, prm)) != MHH_OK) return rc;
    if ((rc = surface_exec_impl<TF>(c, f, prm, s)) != MHH_OK) return rc;
    if ((rc = ghost_all_impl<TF>(c, f, prm)) != MHH_OK) return rc;
    return substep_post_impl<TF>(c, f, prm, substep, dt);
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

// ---- one full RK3 step: eager, or as a replayed CUDA graph --------------------------------------------------------------
inline unsigned long long fnv1a(unsigned long long h, const void* p, size_t n)
{
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

template <typename TF>
int step_eager(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, double dt)
{
    for (int ss = 0; ss < 3; ++ss)
    {
        const int rc = substep_impl<TF>(c, f, prm, ss, dt);
        if (rc != MHH_OK) return rc;
    }
    return MHH_OK;
}

template <typename TF>
int step_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, double dt)
{
    if (!c->use_graph || c->nranks > 1 || c->desc.npy > 1 || c->prof || c->graph_failed || !c->own_stream
        || (c->graph_mode < 0 && (size_t)c->g.ncells * sizeof(TF) > ((size_t)512 << 20)))
        return step_eager<TF>(c, f, prm, dt);
    // everything a captured launch bakes in: the field / parameter structs (pointers, switches, viscosities), dt, the registered
    // forcing and closure.  The user's stream is not part of it: the graph always runs on the context's own stream.
    unsigned long long key = 1469598103934665603ull;
    key = fnv1a(key, f, sizeof(*f)); key = fnv1a(key, prm, sizeof(*prm)); key = fnv1a(key, &dt, sizeof(dt));
    const int flags[5] = {c->forcing_set ? 1 : 0, c->tke2_set ? 1 : 0, c->overlap ? 1 : 0, c->buoy_set ? 1 : 0, c->moist_set ? 1 : 0};
    key = fnv1a(key, flags, sizeof(flags));
    if (c->buoy_set) key = fnv1a(key, &c->buoy, sizeof(c->buoy));
    if (c->moist_set) key = fnv1a(key, &c->moist, sizeof(c->moist));
    if (c->forcing_set) key = fnv1a(key, &c->forcing, sizeof(c->forcing));
    if (c->tke2_set) key = fnv1a(key, &c->tke2, sizeof(c->tke2));
    if (key == 0) key = 1;
    cudaStream_t user = c->stream;
    if (!(c->graph_exec && c->graph_key == key))
    {
        if (c->graph_seen != key) { c->graph_seen = key; return step_eager<TF>(c, f, prm, dt); }       // first sighting: lazy set-up runs here
        // second sighting: capture the step on the own stream (the user's stream may be the legacy default stream, which cannot
        // be captured), instantiate, and fall through to the replay
        if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
        if (!c->ev_g0 && (cudaEventCreateWithFlags(&c->ev_g0, cudaEventDisableTiming) != cudaSuccess ||
                          cudaEventCreateWithFlags(&c->ev_g1, cudaEventDisableTiming) != cudaSuccess))
        { cudaGetLastError(); c->graph_failed = true; return step_eager<TF>(c, f, prm, dt); }
        const long long l0 = c->launches;
        if (cudaStreamBeginCapture(c->own_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
        { cudaGetLastError(); c->graph_failed = true; return step_eager<TF>(c, f, prm, dt); }
        c->stream = c->own_stream;
        const int rc = step_eager<TF>(c, f, prm, dt);
        c->stream = user;
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(c->own_stream, &graph);
        const long long nl = c->launches - l0;
        c->launches = l0;
        if (rc != MHH_OK) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); c->graph_failed = true; return rc; }
        if (e != cudaSuccess || !graph || cudaGraphInstantiate(&c->graph_exec, graph, 0) != cudaSuccess)
        {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError(); c->graph_exec = nullptr; c->graph_failed = true;
            return step_eager<TF>(c, f, prm, dt);                                                        // nothing has run yet
        }
        cudaGraphDestroy(graph);
        c->graph_key = key; c->graph_launches = nl;
    }
    // replay: own stream after everything queued on the user's stream, user's stream after the graph
    if (user != c->own_stream)
    {
        CUDA_TRY(c, cudaEventRecord(c->ev_g0, user));
        CUDA_TRY(c, cudaStreamWaitEvent(c->own_stream, c->ev_g0, 0));
    }
    CUDA_TRY(c, cudaGraphLaunch(c->graph_exec, c->own_stream));
    if (user != c->own_stream)
    {
        CUDA_TRY(c, cudaEventRecord(c->ev_g1, c->own_stream));
        CUDA_TRY(c, cudaStreamWaitEvent(user, c->ev_g1, 0));
    }
    c->launches += c->graph_launches;
    c->graph_replays++;
    return MHH_OK;
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
    {
        const int rc = step_impl<TF>(c, f, prm, dt);
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


#define INST(TF) \
    template int slab_barrier<TF>(Ctx<TF>*, const char*); template int exchange_ns<TF>(Ctx<TF>*, TF* const*, int, int, int); \
    template int cyclic_impl<TF>(Ctx<TF>*, TF*, int, bool); template int cyclic_fields<TF>(Ctx<TF>*, TF* const*, int);
INST(double)
INST(float)
#undef INST

} // namespace mhhhost

using namespace mhhhost;


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

// entry points that dispatch on the dtype by hand still have to run on the context's device
#define SET_DEVICE(ctx) do { cudaError_t e_ = cudaSetDevice((ctx)->device); \
    if (e_ != cudaSuccess) { (ctx)->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e_); return MHH_E_CUDA; } } while (0)

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
#ifdef MHH_TEST_KNOBS       // fault injection for tools/gpu_fallback.sh only: build with `make EXTRA=-DMHH_TEST_KNOBS`; not in the shipped library
    { const char* e = getenv("MHH_FAIL_PEER_RANK"); if (e && atoi(e) == ctx->rank) { ctx->err = "open_peers: failure injected by MHH_FAIL_PEER_RANK (test knob)"; return MHH_E_CUDA; } }
#endif
    DISPATCH1(ctx, ([&]() -> int {
        if (c->peers.on) { c->err = "open_peers: already open"; return MHH_E_INVALID; }
        if (c->g.jtot == 1) return MHH_OK;
        const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(all);
        PeerPtrs<TF> pp{};
        void *ps = nullptr, *pn = nullptr;
        // on any failure close what has been opened so far (peers.on stays 0, so nobody else would)
        auto fail = [&](cudaError_t e, const char* what) -> int {
            for (int q = 0; q < c->nranks; ++q)
                if (q != c->rank) { if (pp.x[q]) cudaIpcCloseMemHandle(pp.x[q]); if (pp.y[q]) cudaIpcCloseMemHandle(pp.y[q]); }
            if (ps) cudaIpcCloseMemHandle(ps);
            if (pn && pn != ps) cudaIpcCloseMemHandle(pn);
            c->err = std::string(what) + ": " + cudaGetErrorString(e);
            return MHH_E_CUDA; };
        for (int r = 0; r < c->nranks; ++r)
        {
            if (r == c->rank) { pp.x[r] = c->spec; pp.y[r] = c->specT; continue; }
            void *px = nullptr, *py = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&px, h[3 * r], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return fail(e, "cudaIpcOpenMemHandle(x side)");
            pp.x[r] = static_cast<TF*>(px);
            e = cudaIpcOpenMemHandle(&py, h[3 * r + 1], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return fail(e, "cudaIpcOpenMemHandle(y side)");
            pp.y[r] = static_cast<TF*>(py);
        }
        if (c->phalo && !(getenv("MHH_NO_PEER_HALO") && getenv("MHH_NO_PEER_HALO")[0] == '1'))
        {
            const int south = (c->rank + c->nranks - 1) % c->nranks, north = (c->rank + 1) % c->nranks;
            cudaError_t e = cudaIpcOpenMemHandle(&ps, h[3 * south + 2], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { ps = nullptr; return fail(e, "cudaIpcOpenMemHandle(halo south)"); }
            if (north == south) pn = ps;
            else
            {
                e = cudaIpcOpenMemHandle(&pn, h[3 * north + 2], cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) { pn = nullptr; return fail(e, "cudaIpcOpenMemHandle(halo north)"); }
            }
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

// ---- layout of the fused Pres_2 path (poisson_fused.cuh, struct Spec2) ----
int mhh_slab2_layout(int itot, int jtot, int ktot, int npy, int rank, mhh_slab2_info* out)
{
    if (!out || npy < 1 || rank < 0 || rank >= npy || itot < 2 || jtot < 1 || ktot < 1 || jtot % npy != 0 || itot / 2 + 1 < npy) return MHH_E_INVALID;
    const Spec2 l = make_spec2(itot, jtot, ktot, npy, rank);
    out->nm = l.nm; out->mcl = l.mcl; out->m_off = l.m_off; out->jmax = l.jmax; out->npan = l.npan; out->ksplit = p2_ksplit(ktot);
    out->xside_elems = l.xside_elems(); out->yside_elems = l.yside_elems();
    return MHH_OK;
}

long long mhh_slab2_yindex(int itot, int jtot, int ktot, int npy, int owner, int src, int ml, int k, int jl)
{
    if (npy < 1 || owner < 0 || owner >= npy || src < 0 || src >= npy || jtot % npy != 0 || itot / 2 + 1 < npy) return -1;
    const Spec2 l = make_spec2(itot, jtot, ktot, npy, owner);
    if (ml < 0 || ml >= l.mcl || k < 0 || k >= ktot || jl < 0 || jl >= l.jmax) return -1;
    return l.yidx(l.mcl, src, ml, k, jl);
}

long long mhh_slab2_xindex(int itot, int jtot, int ktot, int npy, int mode_owner, int k, int jl, int ml)
{
    if (npy < 1 || mode_owner < 0 || mode_owner >= npy || jtot % npy != 0 || itot / 2 + 1 < npy) return -1;
    const Spec2 l = make_spec2(itot, jtot, ktot, npy, 0);
    if (ml < 0 || ml >= l.count(mode_owner) || k < 0 || k >= ktot || jl < 0 || jl >= l.jmax) return -1;
    return l.xidx(mode_owner, k, jl / P2_ROWS, ml, jl % P2_ROWS);
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
{ if (ctx) ctx->drop_graph(); DISPATCH1(ctx, set_basestate_impl<TF>(c, rhoref, rhorefh, thref, threfh)); }

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
    if (ctx && swadvec != 25 && swadvec != 2 && swadvec != 4 && swadvec != 41 && swadvec != 24 && swadvec != 262)
    { ctx->err = "advec_exec: swadvec must be 25 (2i5), 2, 24 (2i4), 262 (2i62), 4 or 41 (4m)"; return MHH_E_INVALID; }
    if (swadvec == 24 || swadvec == 262) DISPATCH1(ctx, adv2i_impl<TF>(c, f, swadvec));
    if (swadvec == 2) DISPATCH1(ctx, o2_impl<TF>(c, f, true, false, false));
    if (swadvec == 4 || swadvec == 41) DISPATCH1(ctx, o4_impl<TF>(c, f, swadvec, false));
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
    DISPATCH1(ctx, o4_impl<TF>(c, f, 0, true));
}

int mhh_diff_2_get_dn(mhh_ctx* ctx, const mhh_fields* f, double dt, double* dn)
{
    if (!ctx || !f || !dn) return MHH_E_INVALID;
    SET_DEVICE(ctx);
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
    SET_DEVICE(ctx);
    if (swadvec != 25 && swadvec != 2 && swadvec != 4 && swadvec != 41 && swadvec != 24 && swadvec != 262)
    { ctx->err = "advec_get_cfl: swadvec must be 25 (2i5), 2, 24 (2i4), 262 (2i62), 4 or 41 (4m)"; return MHH_E_INVALID; }
    int rc;
    if (swadvec == 2 || swadvec == 4 || swadvec == 41 || swadvec == 24 || swadvec == 262)
    {
        if (ctx->dtype == MHH_F64) { rc = o2_cfl_impl<double>(static_cast<Ctx<double>*>(ctx), f, cfl, swadvec); if (rc == MHH_OK) *cfl = *cfl * dt; }
        else { rc = o2_cfl_impl<float>(static_cast<Ctx<float>*>(ctx), f, cfl, swadvec); if (rc == MHH_OK) *cfl = (double)((float)*cfl * (float)dt); }
        return rc;
    }
    if (ctx->dtype == MHH_F64) { typedef double TF; rc = reduce_mode_impl<TF>(static_cast<Ctx<TF>*>(ctx), 0, P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), 0, 0, 0, cfl); if (rc == MHH_OK) *cfl = *cfl * dt; }
    else { typedef float TF; rc = reduce_mode_impl<TF>(static_cast<Ctx<TF>*>(ctx), 0, P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), 0, 0, 0, cfl); if (rc == MHH_OK) *cfl = (double)((float)*cfl * (float)dt); }
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
    SET_DEVICE(ctx);
    int rc;
    if (ctx->dtype == MHH_F64)
    {
        typedef double TF; Ctx<TF>* c = static_cast<Ctx<TF>*>(ctx);
        const TF tprfac = TF(1) / std::min(TF(1.), (TF)prm->tPr);
        rc = reduce_mode_impl<TF>(c, 1, P<TF>(f->evisc), nullptr, nullptr, tprfac, (TF)(1. / ((double)c->g.dx * c->g.dx)), (TF)(1. / ((double)c->g.dy * c->g.dy)), dn);
    }
    else
    {
        typedef float TF; Ctx<TF>* c = static_cast<Ctx<TF>*>(ctx);
        const TF tprfac = TF(1) / std::min(TF(1.), (TF)prm->tPr);
        rc = reduce_mode_impl<TF>(c, 1, P<TF>(f->evisc), nullptr, nullptr, tprfac, (TF)(1. / ((double)c->g.dx * c->g.dx)), (TF)(1. / ((double)c->g.dy * c->g.dy)), dn);
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
    SET_DEVICE(ctx);
    if (swpres != 2 && swpres != 4) { ctx->err = "pres_check_divergence: swpres must be 2 or 4"; return MHH_E_INVALID; }
    if (swpres == 4)
    {
        if (ctx->dtype == MHH_F64) return pres4_div_impl<double>(static_cast<Ctx<double>*>(ctx), f, divmax);
        return pres4_div_impl<float>(static_cast<Ctx<float>*>(ctx), f, divmax);
    }
    if (ctx->dtype == MHH_F64) { typedef double TF; Ctx<TF>* c = static_cast<Ctx<TF>*>(ctx); NEED_BASE(c); return reduce_mode_impl<TF>(c, 2, P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), 0, 0, 0, divmax); }
    else { typedef float TF; Ctx<TF>* c = static_cast<Ctx<TF>*>(ctx); NEED_BASE(c); return reduce_mode_impl<TF>(c, 2, P<TF>(f->u), P<TF>(f->v), P<TF>(f->w), 0, 0, 0, divmax); }
}

int mhh_pres_fft_roundtrip(mhh_ctx* ctx, const void* in_compact, void* out_compact, int solve)
{ DISPATCH1(ctx, fft_roundtrip_impl<TF>(c, P<TF>(in_compact), P<TF>(out_compact), solve)); }

int mhh_timeloop_rk3(mhh_ctx* ctx, void* a, void* at, int substep, double dt)
{ DISPATCH1(ctx, rk3_impl<TF>(c, P<TF>(a), P<TF>(at), substep, dt, nullptr)); }

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

int mhh_buffer_exec(mhh_ctx* ctx, const mhh_fields* f, const mhh_forcing* fo)
{
    DISPATCH1(ctx, buffer_exec_impl<TF>(c, f, fo));
}

int mhh_force_exec(mhh_ctx* ctx, const mhh_fields* f, const mhh_forcing* fo, double sub_dt)
{
    DISPATCH1(ctx, force_exec_impl<TF>(c, f, fo, sub_dt));
}

int mhh_dycore_set_forcing(mhh_ctx* ctx, const mhh_forcing* fo)
{
    if (!ctx) return MHH_E_INVALID;
    if (fo) { ctx->forcing = *fo; ctx->forcing_set = true; } else ctx->forcing_set = false;
    ctx->drop_graph();
    return MHH_OK;
}

int mhh_boundary_surface_init(mhh_ctx* ctx, double z0m, double z0h, int mbcbot, int thermobc)
{
    DISPATCH1(ctx, surface_init_impl<TF>(c, z0m, z0h, mbcbot, thermobc));
}

int mhh_boundary_surface_exec(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, const mhh_surface* s)
{
    if (!f || !prm) return MHH_E_INVALID;
    DISPATCH1(ctx, surface_exec_impl<TF>(c, f, prm, s));
}

int mhh_dycore_substep_surface(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, const mhh_surface* s, int substep, double dt)
{
    if (!f || !prm) return MHH_E_INVALID;
    DISPATCH1(ctx, substep_surface_impl<TF>(c, f, prm, s, substep, dt));
}

int mhh_dycore_step(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, double dt)
{
    if (!f || !prm) return MHH_E_INVALID;
    DISPATCH1(ctx, step_impl<TF>(c, f, prm, dt));
}

long long mhh_graph_replays(const mhh_ctx* ctx) { return ctx ? ctx->graph_replays : 0; }

int mhh_dycore_step_host(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, double dt, int nsteps,
                         void* h_u, void* h_v, void* h_w, void* const* h_s)
{
    if (!f || !prm) return MHH_E_INVALID;
    DISPATCH1(ctx, step_host_impl<TF>(c, f, prm, dt, nsteps, h_u, h_v, h_w, h_s));
}

} // extern "C"
