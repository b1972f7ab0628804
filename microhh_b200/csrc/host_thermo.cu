// mhhb200 -- host drivers of Thermo_buoy<TF> (src/thermo_buoy.cxx): exec, get_thermo_field("N2"), registration for the fused
// sub-steps.  (Thermo_dry's buoyancy rides inside the fused tendency kernels, host_tend.cu.)
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

#define INSTANTIATE(TF) \
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

} // extern "C"
