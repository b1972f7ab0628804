// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's dry-thermodynamics kernels: calc_N2 (reference
// src/thermo_dry.cxx:66-78) and calc_buoyancy_tend_2nd (src/thermo_dry.cxx:165-179).
#include <src/thermo_dry.cxx>
#include "ref_common.h"

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_thermo_dry_N2_##SFX(TF* N2, const TF* th, const TF* dzi, const TF* thref) \
{ const Ref_geom& g = ref_geom; \
  calc_N2<TF>(N2, th, dzi, thref, g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells, g.kcells); } \
MHH_EXPORT void ref_thermo_dry_buoyancy_tend_2nd_##SFX(TF* wt, const TF* th, const TF* threfh) \
{ const Ref_geom& g = ref_geom; \
  calc_buoyancy_tend_2nd<TF>(wt, th, threfh, g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells); }

DEFINE(double, f64)
DEFINE(float, f32)
