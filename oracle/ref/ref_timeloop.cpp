// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's low-storage RK3 update `rk3` (reference src/timeloop.cxx:250-286).
#include <src/timeloop.cxx>
#include "ref_common.h"

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_rk3_##SFX(TF* a, TF* at, int substep, TF dt) \
{ const Ref_geom& g = ref_geom; \
  rk3<TF>(a, at, substep, dt, g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells, g.icells*g.jcells*g.kcells); }

DEFINE(double, f64)
DEFINE(float, f32)
