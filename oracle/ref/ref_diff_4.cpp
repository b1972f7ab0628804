// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's Diff_4 CPU kernels (reference src/diff_4.cxx:40-175).
#include <src/diff_4.cxx>
#include "ref_common.h"

#define GEOM const Ref_geom& g = ref_geom
#define RANGE g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_diff_4_c_##SFX(TF* at, const TF* a, TF visc, TF dx, TF dy, const TF* dzi4, const TF* dzhi4) \
{ GEOM; if (g.jtot == 1) diff_c<TF, false>(at, a, visc, RANGE, dx, dy, dzi4, dzhi4); else diff_c<TF, true>(at, a, visc, RANGE, dx, dy, dzi4, dzhi4); } \
MHH_EXPORT void ref_diff_4_w_##SFX(TF* wt, const TF* w, TF visc, TF dx, TF dy, const TF* dzi4, const TF* dzhi4) \
{ GEOM; if (g.jtot == 1) diff_w<TF, false>(wt, w, visc, RANGE, dx, dy, dzi4, dzhi4); else diff_w<TF, true>(wt, w, visc, RANGE, dx, dy, dzi4, dzhi4); }

DEFINE(double, f64)
DEFINE(float, f32)
