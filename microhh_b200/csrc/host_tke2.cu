// mhhb200 -- host drivers of the Deardorff SGS-TKE closure (Diff_tke2<TF>, src/diff_tke2.cxx) and of the Limiter's tendency
// limiter (src/limiter.cxx).  Diff_tke2::exec runs through tend_impl (host_tend.cu) with the per-scalar eddy viscosity.
#include "host_common.cuh"

namespace mhhhost {

template <typename TF>
int tke2_check(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, const mhh_tke2* tke)
{
    NEED(c, f, "fields"); NEED(c, prm, "params"); NEED(c, tke, "tke2");
    if (c->g.dzi4) { c->err = "diff_tke2: second-order grids only (src/diff_tke2.cxx:555)"; return MHH_E_INVALID; }
    if (!prm->surface_model) { c->err = "diff_tke2 requires a surface model (src/diff_tke2.cxx:557)"; return MHH_E_INVALID; }
    if (tke->isgstke < 0 || tke->isgstke >= f->ns) { c->err = "diff_tke2: isgstke is not a scalar of mhh_fields"; return MHH_E_INVALID; }
    if (prm->swthermo != 0) NEED(c, tke->eviscs, "eviscs");
    return MHH_OK;
}

// Diff_tke2::exec_viscosity (src/diff_tke2.cxx:799-983), one kernel + the cyclic fill of evisc (and eviscs)
template <typename TF>
int tke2_visc_impl(Ctx<TF>* c, const mhh_fields* f, const mhh_params* prm, const mhh_tke2* tke, const TF* n2)
{
    NEED_BASE(c);
    int rc = tke2_check<TF>(c, f, prm, tke);
    if (rc != MHH_OK) return rc;
    NEED(c, f->evisc, "evisc"); NEED(c, f->u, "u"); NEED(c, f->v, "v"); NEED(c, f->w, "w");
    NEED(c, f->s[tke->isgstke], "sgstke"); NEED(c, f->st[tke->isgstke], "sgstke tendency");
    NEED(c, f->dudz_mo, "dudz_mo"); NEED(c, f->dvdz_mo, "dvdz_mo"); NEED(c, f->z0m, "z0m");
    Tke2Args<TF> a{};
    a.evisc = P<TF>(f->evisc); a.eviscs = P<TF>(tke->eviscs);
    a.st = P<TF>(f->st[tke->isgstke]); a.e = P<TF>(f->s[tke->isgstke]);
    a.u = P<TF>(f->u); a.v = P<TF>(f->v); a.w = P<TF>(f->w);
    a.dudz = P<TF>(f->dudz_mo); a.dvdz = P<TF>(f->dvdz_mo); a.dbdz = P<TF>(f->dbdz_mo); a.z0m = P<TF>(f->z0m);
    a.cn = (TF)tke->cn; a.cm = (TF)tke->cm; a.ch1 = (TF)tke->ch1; a.ch2 = (TF)tke->ch2; a.ce1 = (TF)tke->ce1; a.ce2 = (TF)tke->ce2;
    a.mason = prm->sw_mason; a.buoy = prm->swthermo != 0;
    if (a.buoy)
    {
        NEED(c, f->dbdz_mo, "dbdz_mo");
        a.n2 = n2; a.n2mode = n2 ? 0 : 1;
        if (!n2)
        {
            if (tke->isgstke == 0 || !f->s[0]) { c->err = "diff_tke2: no N2 field and scalar 0 is not th"; return MHH_E_INVALID; }
            a.th = P<TF>(f->s[0]);
        }
    }
    tke2_visc_kernel<TF><<<c->grd_interior(), c->blk(), 0, c->stream>>>(a, c->g, c->d_mlen0);
    KCHECKN(c, "tke2_visc_kernel");
    if ((rc = cyclic_impl<TF>(c, a.evisc, MHH_EDGE_BOTH, false)) != MHH_OK) return rc;
    if (a.buoy) rc = cyclic_impl<TF>(c, a.eviscs, MHH_EDGE_BOTH, false);
    return rc;
}

template <typename TF>
int limiter_impl(Ctx<TF>* c, TF* at, const TF* a, TF min_value, TF sub_dt)
{
    NEED(c, at, "tendency"); NEED(c, a, "field");
    limiter_kernel<TF><<<c->grd_interior(), c->blk(), 0, c->stream>>>(at, a, min_value, sub_dt, c->g);
    KCHECKN(c, "limiter_kernel");
    return MHH_OK;
}

template <typename TF>
int tke2_create_impl(Ctx<TF>* c, TF* e)
{
    NEED(c, e, "sgstke");
    tke2_min_kernel<TF><<<c->grd_interior(), c->blk(), 0, c->stream>>>(e, c->g);
    KCHECKN(c, "tke2_min_kernel");
    return cyclic_impl<TF>(c, e, MHH_EDGE_BOTH, false);
}

#define INSTANTIATE(TF) \
    template int tke2_check<TF>(Ctx<TF>*, const mhh_fields*, const mhh_params*, const mhh_tke2*); \
    template int tke2_visc_impl<TF>(Ctx<TF>*, const mhh_fields*, const mhh_params*, const mhh_tke2*, const TF*); \
    template int limiter_impl<TF>(Ctx<TF>*, TF*, const TF*, TF, TF);
INSTANTIATE(double)
INSTANTIATE(float)
#undef INSTANTIATE

} // namespace mhhhost

using namespace mhhhost;

#define DISPATCH1(ctx, expr) \
    do { if (!(ctx)) return MHH_E_INVALID; \
         cudaError_t e_ = cudaSetDevice((ctx)->device); \
         if (e_ != cudaSuccess) { (ctx)->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e_); return MHH_E_CUDA; } \
         if ((ctx)->dtype == MHH_F64) { typedef double TF; Ctx<TF>* c = static_cast<Ctx<TF>*>(ctx); (void)c; return (expr); } \
         else { typedef float TF; Ctx<TF>* c = static_cast<Ctx<TF>*>(ctx); (void)c; return (expr); } } while (0)

extern "C" {

int mhh_diff_tke2_create(mhh_ctx* ctx, void* sgstke)
{ DISPATCH1(ctx, tke2_create_impl<TF>(c, P<TF>(sgstke))); }

int mhh_diff_tke2_exec_viscosity(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, const mhh_tke2* tke, const void* n2)
{
    if (!f || !prm || !tke) return MHH_E_INVALID;
    DISPATCH1(ctx, tke2_visc_impl<TF>(c, f, prm, tke, P<TF>(n2)));
}

int mhh_diff_tke2_exec(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, const mhh_tke2* tke)
{
    if (!f || !prm || !tke) return MHH_E_INVALID;
    DISPATCH1(ctx, ([&]() -> int {
        int rc = tke2_check<TF>(c, f, prm, tke);
        if (rc != MHH_OK) return rc;
        return tend_impl<TF>(c, f, prm, false, true, false, tke); })());
}

int mhh_diff_tke2_get_dn(mhh_ctx* ctx, const mhh_fields* f, const mhh_params* prm, const mhh_tke2* tke, double dt, double* dn)
{
    if (!f || !prm || !tke || !dn) return MHH_E_INVALID;
    DISPATCH1(ctx, ([&]() -> int {
        int rc = tke2_check<TF>(c, f, prm, tke);
        if (rc != MHH_OK) return rc;
        NEED(c, f->evisc, "evisc");
        // "When no buoyancy, use eddy viscosity for momentum" (src/diff_tke2.cxx:617-621); tPr_dummy = 1
        const TF* ev = prm->swthermo != 0 ? P<TF>(tke->eviscs) : P<TF>(f->evisc);
        rc = reduce_mode_impl<TF>(c, 1, ev, nullptr, nullptr, TF(1), (TF)(1. / ((double)c->g.dx * c->g.dx)), (TF)(1. / ((double)c->g.dy * c->g.dy)), dn);
        if (rc == MHH_OK) *dn = *dn * dt;
        return rc; })());
}

int mhh_limiter_exec(mhh_ctx* ctx, void* at, const void* a, double min_value, double sub_dt)
{ DISPATCH1(ctx, limiter_impl<TF>(c, P<TF>(at), P<TF>(a), (TF)min_value, (TF)sub_dt)); }

int mhh_dycore_set_tke2(mhh_ctx* ctx, const mhh_tke2* tke)
{
    if (!ctx) return MHH_E_INVALID;
    if (tke) { ctx->tke2 = *tke; ctx->tke2_set = true; }
    else { ctx->tke2 = mhh_tke2{}; ctx->tke2_set = false; }
    ctx->drop_graph();
    return MHH_OK;
}

} // extern "C"
