// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's Diff_2 CPU kernels (reference src/diff_2.cxx:38-86),
// reached by including that translation unit in place.
#include <src/diff_2.cxx>
#include "ref_common.h"

#define GEOM const Ref_geom& g = ref_geom
#define RANGE g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_diff_2_c_##SFX(TF* at, const TF* a, TF visc, TF dx, TF dy, const TF* dzi, const TF* dzhi) \
{ GEOM; diff_c<TF>(at, a, visc, RANGE, dx, dy, dzi, dzhi); } \
MHH_EXPORT void ref_diff_2_w_##SFX(TF* wt, const TF* w, TF visc, TF dx, TF dy, const TF* dzi, const TF* dzhi) \
{ GEOM; diff_w<TF>(wt, w, visc, RANGE, dx, dy, dzi, dzhi); }

DEFINE(double, f64)
DEFINE(float, f32)
