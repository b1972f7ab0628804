// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's Advec_2i5 CPU kernels (reference src/advec_2i5.cxx:59-728) and the flux-limited
// scalar advection it calls for `fluxlimit_list` scalars (reference include/advec_monotonic.h:98-202),
// reached by including that translation unit in place.
#include <src/advec_2i5.cxx>
#include "ref_common.h"

#define GEOM const Ref_geom& g = ref_geom
#define RANGE g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_advec_2i5_u_##SFX(TF* ut, const TF* u, const TF* v, const TF* w, const TF* dzi, TF dx, TF dy, const TF* rhoref, const TF* rhorefh) \
{ GEOM; advec_u<TF>(ut, u, v, w, dzi, dx, dy, rhoref, rhorefh, RANGE); } \
MHH_EXPORT void ref_advec_2i5_v_##SFX(TF* vt, const TF* u, const TF* v, const TF* w, const TF* dzi, TF dx, TF dy, const TF* rhoref, const TF* rhorefh) \
{ GEOM; advec_v<TF>(vt, u, v, w, dzi, dx, dy, rhoref, rhorefh, RANGE); } \
MHH_EXPORT void ref_advec_2i5_w_##SFX(TF* wt, const TF* u, const TF* v, const TF* w, const TF* dzhi, TF dx, TF dy, const TF* rhoref, const TF* rhorefh) \
{ GEOM; advec_w<TF>(wt, u, v, w, dzhi, dx, dy, rhoref, rhorefh, RANGE); } \
MHH_EXPORT void ref_advec_2i5_s_##SFX(TF* st, const TF* s, const TF* u, const TF* v, const TF* w, const TF* dzi, TF dx, TF dy, const TF* rhoref, const TF* rhorefh) \
{ GEOM; advec_s<TF>(st, s, u, v, w, dzi, dx, dy, rhoref, rhorefh, RANGE); } \
MHH_EXPORT void ref_advec_s_lim_##SFX(TF* st, const TF* s, const TF* u, const TF* v, const TF* w, const TF* dzi, TF dx, TF dy, const TF* rhoref, const TF* rhorefh) \
{ GEOM; Advec_monotonic::advec_s_lim<TF>(st, s, u, v, w, dzi, dx, dy, rhoref, rhorefh, RANGE); } \
MHH_EXPORT double ref_advec_2i5_cfl_##SFX(const TF* u, const TF* v, const TF* w, const TF* dzi, TF dx, TF dy, TF dt) \
{ GEOM; alignas(16) static char mbuf[sizeof(Master)]; \
  return (double)calc_cfl<TF>(u, v, w, dzi, dx, dy, dt, *reinterpret_cast<Master*>(mbuf), RANGE); }

DEFINE(double, f64)
DEFINE(float, f32)
