// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's Advec_2i4 CPU kernels (reference src/advec_2i4.cxx:53-518), reached by including that
// translation unit in place; exec passes gd.dxi = 1./gd.dx (src/grid.cxx:252-253, a double division narrowed to TF).
#include <src/advec_2i4.cxx>
#include "ref_common.h"

#define GEOM const Ref_geom& g = ref_geom
#define RANGE g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_advec_2i4_u_##SFX(TF* ut, const TF* u, const TF* v, const TF* w, const TF* dzi, TF dx, TF dy, const TF* rhoref, const TF* rhorefh) \
{ GEOM; advec_u<TF>(ut, u, v, w, dzi, (TF)(1./dx), (TF)(1./dy), rhoref, rhorefh, RANGE); } \
MHH_EXPORT void ref_advec_2i4_v_##SFX(TF* vt, const TF* u, const TF* v, const TF* w, const TF* dzi, TF dx, TF dy, const TF* rhoref, const TF* rhorefh) \
{ GEOM; advec_v<TF>(vt, u, v, w, dzi, (TF)(1./dx), (TF)(1./dy), rhoref, rhorefh, RANGE); } \
MHH_EXPORT void ref_advec_2i4_w_##SFX(TF* wt, const TF* u, const TF* v, const TF* w, const TF* dzhi, TF dx, TF dy, const TF* rhoref, const TF* rhorefh) \
{ GEOM; advec_w<TF>(wt, u, v, w, dzhi, (TF)(1./dx), (TF)(1./dy), rhoref, rhorefh, RANGE); } \
MHH_EXPORT void ref_advec_2i4_s_##SFX(TF* st, const TF* s, const TF* u, const TF* v, const TF* w, const TF* dzi, TF dx, TF dy, const TF* rhoref, const TF* rhorefh) \
{ GEOM; advec_s<TF>(st, s, u, v, w, dzi, (TF)(1./dx), (TF)(1./dy), rhoref, rhorefh, RANGE); } \
MHH_EXPORT double ref_advec_2i4_cfl_##SFX(const TF* u, const TF* v, const TF* w, const TF* dzi, TF dx, TF dy, TF dt) \
{ GEOM; alignas(16) static char mbuf[sizeof(Master)]; \
  return (double)calc_cfl<TF>(u, v, w, dzi, (TF)(1./dx), (TF)(1./dy), dt, *reinterpret_cast<Master*>(mbuf), RANGE); }

DEFINE(double, f64)
DEFINE(float, f32)
