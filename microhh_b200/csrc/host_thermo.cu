// mhhb200 -- host drivers of Thermo_buoy<TF> (src/thermo_buoy.cxx) and Thermo_moist<TF> (src/thermo_moist.cxx): exec, the
// get_thermo_field diagnostics, the moist base state, registration for the fused sub-steps.  (Thermo_dry's buoyancy rides
// inside the fused tendency kernels, host_tend.cu.)
#include "host_common.cuh"

namespace mhhhost {

// Thermo_buoy::exec (src/thermo_buoy.cxx:345-391)
template <typename TF>
int thermo_buoy_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_thermo_buoy* tb)
{
    NEED(c, f, "fields"); NEED(c, tb, "thermo_buoy");
    const GridDev<TF>& g = c->g;
    if (f->ns < 1) { c->err = "thermo_buoy: scalar 0 must be the buoyancy b"; return MHH_E_INVALID; }
    NEED(c, f->s[0], "b"); NEED(c, f->wt, "wt");
    BuoyArgs<TF> a{};
    // `bs.alpha` and `bs.n2` are TF members; has_slope / has_N2 test them after the narrowing (src/thermo_buoy.cxx:319-324)
    const TF alpha = (TF)tb->alpha, n2 = (TF)tb->n2;
    a.slope = (std::abs(alpha) > TF(0.) || std::abs(n2) > TF(0.)) ? 1 : 0;
    a.baroclinic = tb->swbaroclinic ? 1 : 0;
    a.ut = P<TF>(f->ut); a.wt = P<TF>(f->wt); a.bt = P<TF>(f->st[0]);
    a.b = P<TF>(f->s[0]); a.u = P<TF>(f->u); a.v = P<TF>(f->v); a.w = P<TF>(f->w);
    if (a.slope) { NEED(c, a.ut, "ut"); NEED(c, a.bt, "b tendency"); NEED(c, a.u, "u"); NEED(c, a.w, "w"); }
    if (a.baroclinic) { NEED(c, a.bt, "b tendency"); NEED(c, a.v, "v"); }
    a.sinalpha = std::sin(alpha); a.cosalpha = std::cos(alpha);       // std::sin(TF) on the host, like the reference's kernels
    a.n2 = n2; a.utrans = (TF)tb->utrans; a.dbdy_ls = (TF)tb->dbdy_ls;
    dim3 b(64, 4), gr((g.imax + 63) / 64, (g.jmax + 3) / 4, g.kmax);
    if (g.dzi4) thermo_buoy_kernel<TF, 4><<<gr, b, 0, c->stream>>>(a, g);
    else        thermo_buoy_kernel<TF, 2><<<gr, b, 0, c->stream>>>(a, g);
    KCHECKN(c, "thermo_buoy_kernel");
    return MHH_OK;
}

template <typename TF>
int thermo_buoy_n2_impl(Ctx<TF>* c, TF* n2, const TF* b, double bg_n2)
{
    NEED(c, n2, "n2"); NEED(c, b, "b");
    const GridDev<TF>& g = c->g;
    dim3 bl(64, 4), gr((g.imax + 63) / 64, (g.jmax + 3) / 4, g.kmax);
    thermo_buoy_n2_kernel<TF><<<gr, bl, 0, c->stream>>>(n2, b, (TF)bg_n2, g);
    KCHECKN(c, "thermo_buoy_n2_kernel");
    return MHH_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Thermo_moist (src/thermo_moist.cxx)
// ---------------------------------------------------------------------------------------------------------------------------
template <typename TF>
int moist_alloc(Ctx<TF>* c)
{
    if (c->d_moist) return MHH_OK;
    const size_t kc = c->g.kcells;
    CUDA_TRY(c, cudaMalloc(&c->d_moist, sizeof(TF) * kc * 10));
    CUDA_TRY(c, cudaMemset(c->d_moist, 0, sizeof(TF) * kc * 10));
    CUDA_TRY(c, cudaMalloc(&c->d_moist_flag, 2 * sizeof(int)));        // [0] non-converged adjustments, [1] sweeps of the last base state
    CUDA_TRY(c, cudaMemset(c->d_moist_flag, 0, 2 * sizeof(int)));
    return MHH_OK;
}

template <typename TF>
int moist_check(Ctx<TF>* c, const mhh_fields* f, const mhh_thermo_moist* tm)
{
    NEED(c, f, "fields"); NEED(c, tm, "thermo_moist");
    if (c->g.dzi4) { c->err = "thermo_moist: second-order grids only (Thermo_moist::exec has no 4th-order branch)"; return MHH_E_INVALID; }
    if (tm->ithl < 0 || tm->ithl >= f->ns || tm->iqt < 0 || tm->iqt >= f->ns || tm->ithl == tm->iqt)
    { c->err = "thermo_moist: ithl / iqt are not two scalars of mhh_fields"; return MHH_E_INVALID; }
    NEED(c, f->s[tm->ithl], "thl"); NEED(c, f->s[tm->iqt], "qt");
    if (!c->moist_profiles_set) { c->err = "thermo_moist: no base state (mhh_thermo_moist_calc_base_state / _set_profiles)"; return MHH_E_INVALID; }
    return MHH_OK;
}

// calc_base_state on the device from device mean profiles
template <typename TF>
int moist_base_state_launch(Ctx<TF>* c, const TF* thlmean, const TF* qtmean, double pbot, bool cold)
{
    TF* scratch = c->d_moist + 8 * (size_t)c->g.kcells;
    moist_base_state_kernel<TF><<<1, 256, 0, c->stream>>>(c->moist_profiles(), thlmean, qtmean, (TF)pbot, c->g, scratch, scratch + c->g.kcells,
                                                          cold ? 1 : 0, c->d_moist_flag);
    KCHECKN(c, "moist_base_state_kernel");
    return MHH_OK;
}

template <typename TF>
int moist_calc_base_state_impl(Ctx<TF>* c, const TF* thl0, const TF* qt0, double pbot)
{
    NEED(c, thl0, "thl0"); NEED(c, qt0, "qt0");
    int rc = moist_alloc<TF>(c);
    if (rc != MHH_OK) return rc;
    const size_t kc = c->g.kcells;
    TF* means = c->d_moist + 6 * kc;
    CUDA_TRY(c, cudaMemcpyAsync(means, thl0, sizeof(TF) * kc, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(means + kc, qt0, sizeof(TF) * kc, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));          // the host arrays may be pageable and go out of scope
    if ((rc = moist_base_state_launch<TF>(c, means, means + kc, pbot, true)) != MHH_OK) return rc;
    c->moist_profiles_set = true;
    return MHH_OK;
}

template <typename TF>
int moist_profiles_copy(Ctx<TF>* c, void* const* host, bool to_host)
{
    int rc = moist_alloc<TF>(c);
    if (rc != MHH_OK) return rc;
    // a pressure profile travels with its exner function (the kernels read exnref / exnrefh instead of re-evaluating the pow)
    if (!to_host && ((host[0] != nullptr) != (host[6] != nullptr) || (host[1] != nullptr) != (host[7] != nullptr)))
    { c->err = "thermo_moist_set_profiles: pref goes with exnref and prefh with exnrefh (both or neither)"; return MHH_E_INVALID; }
    const MoistProfiles<TF> b = c->moist_profiles();
    TF* dev[8] = {b.pref, b.prefh, b.rho, b.rhoh, b.thv, b.thvh, b.ex, b.exh};
    const size_t bytes = sizeof(TF) * c->g.kcells;
    for (int n = 0; n < 8; ++n)
    {
        if (!host[n]) continue;
        if (to_host) CUDA_TRY(c, cudaMemcpyAsync(host[n], dev[n], bytes, cudaMemcpyDeviceToHost, c->stream));
        else         CUDA_TRY(c, cudaMemcpyAsync(dev[n], host[n], bytes, cudaMemcpyHostToDevice, c->stream));
    }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (!to_host) c->moist_profiles_set = true;
    return MHH_OK;
}

// Thermo_moist::exec (src/thermo_moist.cxx:1415-1447)
template <typename TF>
int thermo_moist_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_thermo_moist* tm)
{
    int rc = moist_check<TF>(c, f, tm);
    if (rc != MHH_OK) return rc;
    NEED(c, f->wt, "wt");
    const GridDev<TF>& g = c->g;
    const TF* thl = P<TF>(f->s[tm->ithl]);
    const TF* qt = P<TF>(f->s[tm->iqt]);
    if (tm->swupdatebasestate)
    {
        // Fields::exec's mean profiles (src/fields.cxx:542-551) + calc_base_state, on the device
        TF* means = c->d_moist + 6 * (size_t)g.kcells;
        moist_mean_profile_kernel<TF><<<dim3(g.kcells, 2), 1024, 0, c->stream>>>(thl, qt, means, means + g.kcells, g, (double)g.itot * (double)g.jtot);
        KCHECKN(c, "moist_mean_profile_kernel");
        if (c->nranks > 1)
        {
            // y slabs: every rank holds sum(local rows) / (itot * jtot); `master.sum(prof, kcells)` of the reference
            // (src/field3d_operators.cxx:65) is the all-reduce of those TF partial means.  Every rank then integrates the same
            // base state.  (Same pattern as the fixed-mass-flux sums of Force; NOT yet run on a multi-GPU box.)
            if (!c->comm) { c->err = "slab context without communicator: call mhh_comm_init first"; return MHH_E_INVALID; }
            NcclApi* api = nccl_api(c->err);
            if (!api) return MHH_E_CUDA;
            NCCL_TRY(c, api, api->AllReduce(means, means, 2 * (size_t)g.kcells, sizeof(TF) == 8 ? ncclFloat64 : ncclFloat32, ncclSum, c->comm, c->stream));
        }
        if ((rc = moist_base_state_launch<TF>(c, means, means + g.kcells, tm->pbot, false)) != MHH_OK) return rc;   // starts from the previous base state
    }
    if (g.kmax > 1)
    {
        constexpr int KCH = 8;          // levels per thread
        dim3 b(64, 4), gr((g.imax + 63) / 64, (g.jmax + 3) / 4, (g.kmax - 1 + KCH - 1) / KCH);
        moist_buoyancy_tend_kernel<TF, KCH><<<gr, b, 0, c->stream>>>(P<TF>(f->wt), thl, qt, c->moist_profiles().prefh, c->moist_profiles().exh, g.threfh, g, c->d_moist_flag);
        KCHECKN(c, "moist_buoyancy_tend_kernel");
    }
    return MHH_OK;
}

template <typename TF>
int moist_field_impl(Ctx<TF>* c, int which, TF* out, const mhh_fields* f, const mhh_thermo_moist* tm)
{
    int rc = moist_check<TF>(c, f, tm);
    if (rc != MHH_OK) return rc;
    NEED(c, out, "out");
    const GridDev<TF>& g = c->g;
    const TF* thl = P<TF>(f->s[tm->ithl]);
    const TF* qt = P<TF>(f->s[tm->iqt]);
    const MoistProfiles<TF> b = c->moist_profiles();
    dim3 bl(64, 4), gr((g.imax + 63) / 64, (g.jmax + 3) / 4, which == MHH_MOIST_B ? g.kcells : g.kmax);
    if (which == MHH_MOIST_B)       moist_field_kernel<TF, 0><<<gr, bl, 0, c->stream>>>(out, thl, qt, b.pref, b.thv, g, c->d_moist_flag);
    else if (which == MHH_MOIST_QL) moist_field_kernel<TF, 1><<<gr, bl, 0, c->stream>>>(out, thl, qt, b.pref, b.thv, g, c->d_moist_flag);
    else if (which == MHH_MOIST_N2) moist_field_kernel<TF, 2><<<gr, bl, 0, c->stream>>>(out, thl, qt, b.pref, b.thv, g, c->d_moist_flag);
    else { c->err = "thermo_moist_get_thermo_field: which must be MHH_MOIST_B, _QL or _N2"; return MHH_E_INVALID; }
    KCHECKN(c, "moist_field_kernel");
    return MHH_OK;
}

template <typename TF>
int moist_surf_impl(Ctx<TF>* c, TF* b3d, TF* plane, const mhh_fields* f, const mhh_thermo_moist* tm, int mode)
{
    int rc = moist_check<TF>(c, f, tm);
    if (rc != MHH_OK) return rc;
    NEED(c, plane, mode == 0 ? "bbot" : "bfluxbot");
    const GridDev<TF>& g = c->g;
    const TF* thl2 = P<TF>(mode == 0 ? f->s_bot[tm->ithl] : f->s_fluxbot[tm->ithl]);
    const TF* qt2  = P<TF>(mode == 0 ? f->s_bot[tm->iqt]  : f->s_fluxbot[tm->iqt]);
    if (mode == 0) NEED(c, b3d, "b");
    NEED(c, thl2, mode == 0 ? "s_bot[thl]" : "s_fluxbot[thl]"); NEED(c, qt2, mode == 0 ? "s_bot[qt]" : "s_fluxbot[qt]");
    dim3 bl(64, 4), gr((g.icells + 63) / 64, (g.jcells + 3) / 4);
    moist_surf_kernel<TF><<<gr, bl, 0, c->stream>>>(b3d, plane, P<TF>(f->s[tm->ithl]), thl2, P<TF>(f->s[tm->iqt]), qt2, g.thref, g.threfh, g, mode);
    KCHECKN(c, "moist_surf_kernel");
    return MHH_OK;
}

template <typename TF>
int moist_sweeps_impl(Ctx<TF>* c, int* sweeps)
{
    NEED(c, sweeps, "sweeps");
    *sweeps = 0;
    if (!c->d_moist_flag) return MHH_OK;
    CUDA_TRY(c, cudaMemcpyAsync(sweeps, c->d_moist_flag + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return MHH_OK;
}

template <typename TF>
int moist_nonconverged_impl(Ctx<TF>* c, long long* count)
{
    NEED(c, count, "count");
    *count = 0;
    if (!c->d_moist_flag) return MHH_OK;
    int n = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&n, c->d_moist_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->d_moist_flag, 0, sizeof(int), c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *count = n;
    return MHH_OK;
}

#define INSTANTIATE(TF) \
    template int thermo_moist_impl<TF>(Ctx<TF>*, const mhh_fields*, const mhh_thermo_moist*); \
    template int thermo_buoy_impl<TF>(Ctx<TF>*, const mhh_fields*, const mhh_thermo_buoy*);
INSTANTIATE(double)
INSTANTIATE(float)

} // namespace mhhhost

using namespace mhhhost;

#define DISPATCH1(ctx, expr) do { if (!(ctx)) return MHH_E_INVALID; \
         cudaError_t e_ = cudaSetDevice((ctx)->device); \
         if (e_ != cudaSuccess) { (ctx)->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e_); return MHH_E_CUDA; } \
         if ((ctx)->dtype == MHH_F64) { typedef double TF; Ctx<TF>* c = static_cast<Ctx<TF>*>(ctx); (void)c; return (expr); } \
         else { typedef float TF; Ctx<TF>* c = static_cast<Ctx<TF>*>(ctx); (void)c; return (expr); } } while (0)

extern "C" {

int mhh_thermo_buoy_exec(mhh_ctx* ctx, const mhh_fields* f, const mhh_thermo_buoy* tb)
{ DISPATCH1(ctx, thermo_buoy_impl<TF>(c, f, tb)); }

int mhh_thermo_buoy_n2(mhh_ctx* ctx, void* n2, const void* b, double bg_n2)
{ DISPATCH1(ctx, thermo_buoy_n2_impl<TF>(c, P<TF>(n2), P<TF>(b), bg_n2)); }

int mhh_dycore_set_thermo_buoy(mhh_ctx* ctx, const mhh_thermo_buoy* tb)
{
    if (!ctx) return MHH_E_INVALID;
    std::memset(&ctx->buoy, 0, sizeof(ctx->buoy));          // member-wise copy below: the padding bytes stay zero (the struct is part of the graph key)
    ctx->buoy_set = tb != nullptr;
    if (tb)
    {
        ctx->buoy.alpha = tb->alpha; ctx->buoy.n2 = tb->n2; ctx->buoy.utrans = tb->utrans;
        ctx->buoy.swbaroclinic = tb->swbaroclinic; ctx->buoy.dbdy_ls = tb->dbdy_ls;
    }
    ctx->drop_graph();
    return MHH_OK;
}

int mhh_thermo_moist_calc_base_state(mhh_ctx* ctx, const void* thl0, const void* qt0, double pbot)
{ if (ctx) ctx->drop_graph(); DISPATCH1(ctx, moist_calc_base_state_impl<TF>(c, P<TF>(thl0), P<TF>(qt0), pbot)); }

int mhh_thermo_moist_set_profiles(mhh_ctx* ctx, const void* pref, const void* prefh, const void* rhoref, const void* rhorefh,
                                  const void* thvref, const void* thvrefh, const void* exnref, const void* exnrefh)
{
    void* const h[8] = {const_cast<void*>(pref), const_cast<void*>(prefh), const_cast<void*>(rhoref), const_cast<void*>(rhorefh),
                        const_cast<void*>(thvref), const_cast<void*>(thvrefh), const_cast<void*>(exnref), const_cast<void*>(exnrefh)};
    DISPATCH1(ctx, moist_profiles_copy<TF>(c, h, false));
}

int mhh_thermo_moist_get_profiles(mhh_ctx* ctx, void* pref, void* prefh, void* rhoref, void* rhorefh,
                                  void* thvref, void* thvrefh, void* exnref, void* exnrefh)
{
    void* const h[8] = {pref, prefh, rhoref, rhorefh, thvref, thvrefh, exnref, exnrefh};
    DISPATCH1(ctx, moist_profiles_copy<TF>(c, h, true));
}

int mhh_thermo_moist_exec(mhh_ctx* ctx, const mhh_fields* f, const mhh_thermo_moist* tm)
{ DISPATCH1(ctx, thermo_moist_impl<TF>(c, f, tm)); }

int mhh_thermo_moist_get_thermo_field(mhh_ctx* ctx, int which, void* out, const mhh_fields* f, const mhh_thermo_moist* tm)
{ DISPATCH1(ctx, moist_field_impl<TF>(c, which, P<TF>(out), f, tm)); }

int mhh_thermo_moist_get_buoyancy_surf(mhh_ctx* ctx, void* b, void* bbot, const mhh_fields* f, const mhh_thermo_moist* tm)
{ DISPATCH1(ctx, moist_surf_impl<TF>(c, P<TF>(b), P<TF>(bbot), f, tm, 0)); }

int mhh_thermo_moist_get_buoyancy_fluxbot(mhh_ctx* ctx, void* bfluxbot, const mhh_fields* f, const mhh_thermo_moist* tm)
{ DISPATCH1(ctx, moist_surf_impl<TF>(c, (TF*)nullptr, P<TF>(bfluxbot), f, tm, 1)); }

int mhh_thermo_moist_base_state_sweeps(mhh_ctx* ctx, int* sweeps)
{ DISPATCH1(ctx, moist_sweeps_impl<TF>(c, sweeps)); }

int mhh_thermo_moist_nonconverged(mhh_ctx* ctx, long long* count)
{ DISPATCH1(ctx, moist_nonconverged_impl<TF>(c, count)); }

int mhh_dycore_set_thermo_moist(mhh_ctx* ctx, const mhh_thermo_moist* tm)
{
    if (!ctx) return MHH_E_INVALID;
    std::memset(&ctx->moist, 0, sizeof(ctx->moist));
    ctx->moist_set = tm != nullptr;
    if (tm)
    {
        ctx->moist.ithl = tm->ithl; ctx->moist.iqt = tm->iqt; ctx->moist.pbot = tm->pbot;
        ctx->moist.swupdatebasestate = tm->swupdatebasestate;
    }
    ctx->drop_graph();
    return MHH_OK;
}

} // extern "C"
