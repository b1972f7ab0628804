// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's fully conservative 4th-order advection, Advec_4m (reference src/advec_4m.cxx:50-480; moser180
// as shipped is swadvec=4m).
#include <src/advec_4m.cxx>
#include "ref_common.h"

#define GEOM const Ref_geom& g = ref_geom
#define RANGE g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_advec_4m_u_##SFX(TF* ut, const TF* u, const TF* v, const TF* w, const TF* dzi4, TF dx, TF dy) \
{ GEOM; if (g.jtot == 1) advec_u<TF, false>(ut, u, v, w, dzi4, dx, dy, RANGE); else advec_u<TF, true>(ut, u, v, w, dzi4, dx, dy, RANGE); } \
MHH_EXPORT void ref_advec_4m_v_##SFX(TF* vt, const TF* u, const TF* v, const TF* w, const TF* dzi4, TF dx, TF dy) \
{ GEOM; if (g.jtot == 1) advec_v<TF, false>(vt, u, v, w, dzi4, dx, dy, RANGE); else advec_v<TF, true>(vt, u, v, w, dzi4, dx, dy, RANGE); } \
MHH_EXPORT void ref_advec_4m_w_##SFX(TF* wt, const TF* u, const TF* v, TF* w, const TF* dzhi4, TF dx, TF dy) \
{ GEOM; if (g.jtot == 1) advec_w<TF, false>(wt, u, v, w, dzhi4, dx, dy, RANGE); else advec_w<TF, true>(wt, u, v, w, dzhi4, dx, dy, RANGE); } \
MHH_EXPORT void ref_advec_4m_s_##SFX(TF* st, const TF* s, const TF* u, const TF* v, const TF* w, const TF* dzi4, TF dx, TF dy) \
{ GEOM; if (g.jtot == 1) advec_s<TF, false>(st, s, u, v, w, dzi4, dx, dy, RANGE); else advec_s<TF, true>(st, s, u, v, w, dzi4, dx, dy, RANGE); } \
MHH_EXPORT double ref_advec_4m_cfl_##SFX(const TF* u, const TF* v, const TF* w, const TF* dzi, TF dx, TF dy, TF dt) \
{ GEOM; alignas(16) static char mbuf[sizeof(Master)]; \
  return (double)calc_cfl<TF>(u, v, w, dzi, dx, dy, dt, *reinterpret_cast<Master*>(mbuf), RANGE); }

DEFINE(double, f64)
DEFINE(float, f32)
