// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's Advec_2i62 CPU kernels (reference src/advec_2i62.cxx:59-306), reached by including that
// translation unit in place; the kernels take dx, dy and form their own reciprocals.
#include <src/advec_2i62.cxx>
#include "ref_common.h"

#define GEOM const Ref_geom& g = ref_geom
#define RANGE g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_advec_2i62_u_##SFX(TF* ut, const TF* u, const TF* v, const TF* w, const TF* dzi, TF dx, TF dy, const TF* rhoref, const TF* rhorefh) \
{ GEOM; advec_u<TF>(ut, u, v, w, dzi, dx, dy, rhoref, rhorefh, RANGE); } \
MHH_EXPORT void ref_advec_2i62_v_##SFX(TF* vt, const TF* u, const TF* v, const TF* w, const TF* dzi, TF dx, TF dy, const TF* rhoref, const TF* rhorefh) \
{ GEOM; advec_v<TF>(vt, u, v, w, dzi, dx, dy, rhoref, rhorefh, RANGE); } \
MHH_EXPORT void ref_advec_2i62_w_##SFX(TF* wt, const TF* u, const TF* v, const TF* w, const TF* dzhi, TF dx, TF dy, const TF* rhoref, const TF* rhorefh) \
{ GEOM; advec_w<TF>(wt, u, v, w, dzhi, dx, dy, rhoref, rhorefh, RANGE); } \
MHH_EXPORT void ref_advec_2i62_s_##SFX(TF* st, const TF* s, const TF* u, const TF* v, const TF* w, const TF* dzi, TF dx, TF dy, const TF* rhoref, const TF* rhorefh) \
{ GEOM; advec_s<TF>(st, s, u, v, w, dzi, dx, dy, rhoref, rhorefh, RANGE); } \
MHH_EXPORT double ref_advec_2i62_cfl_##SFX(const TF* u, const TF* v, const TF* w, const TF* dzi, TF dx, TF dy, TF dt) \
{ GEOM; alignas(16) static char mbuf[sizeof(Master)]; \
  return (double)calc_cfl<TF>(u, v, w, dzi, dx, dy, dt, *reinterpret_cast<Master*>(mbuf), RANGE); }

DEFINE(double, f64)
DEFINE(float, f32)
