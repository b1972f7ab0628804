// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's moist-thermodynamics CPU kernels (src/thermo_moist.cxx, include/thermo_moist_functions.h):
// calc_base_state (functions.h:271-340), calc_top_and_bot (:57-76), calc_buoyancy_tend_2nd with the saturation adjustment
// (:77-120; sat_adjust functions.h:164-268), calc_buoyancy / calc_liquid_water (get_thermo_field "b" / "ql", :122-168,
// 230-250), calc_N2 (:459-475), calc_buoyancy_bot / calc_buoyancy_fluxbot (:637-693), and Field3d_operators' mean profile
// (src/field3d_operators.cxx:45-66, restated here as the same loop: the member function needs a live Grid / Master).
#include <src/thermo_moist.cxx>
#include "ref_common.h"
#include <vector>

#define GEOM const Ref_geom& g = ref_geom
#define RANGE g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_moist_base_state_##SFX(TF* pref, TF* prefh, TF* rho, TF* rhoh, TF* thv, TF* thvh, TF* ex, TF* exh, \
        const TF* thlmean, const TF* qtmean, TF pbot, const TF* z, const TF* dz, const TF* dzh) \
{ GEOM; Thermo_moist_functions::calc_base_state<TF>(pref, prefh, rho, rhoh, thv, thvh, ex, exh, thlmean, qtmean, pbot, g.kstart, g.kend, z, dz, dzh); } \
MHH_EXPORT void ref_moist_top_and_bot_##SFX(TF* thl0, TF* qt0, const TF* z, const TF* zh, const TF* dzhi) \
{ GEOM; calc_top_and_bot<TF>(thl0, qt0, z, zh, dzhi, g.kstart, g.kend); } \
MHH_EXPORT void ref_moist_buoyancy_tend_2nd_##SFX(TF* wt, TF* thl, TF* qt, TF* ph, TF* thvrefh) \
{ GEOM; std::vector<TF> tmp(4*(size_t)g.icells*g.jcells); const size_t ij = (size_t)g.icells*g.jcells; \
  calc_buoyancy_tend_2nd<TF>(wt, thl, qt, ph, &tmp[0], &tmp[ij], &tmp[2*ij], &tmp[3*ij], thvrefh, RANGE); } \
MHH_EXPORT void ref_moist_buoyancy_##SFX(TF* b, TF* thl, TF* qt, TF* p, TF* thvref) \
{ GEOM; const size_t nc = (size_t)g.icells*g.jcells*g.kcells; std::vector<TF> tmp(2*nc); \
  calc_buoyancy<TF>(b, thl, qt, p, &tmp[0], &tmp[nc], thvref, g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.kcells, g.icells, g.icells*g.jcells); } \
MHH_EXPORT void ref_moist_liquid_water_##SFX(TF* ql, TF* thl, TF* qt, TF* p) \
{ GEOM; calc_liquid_water<TF>(ql, thl, qt, p, RANGE); } \
MHH_EXPORT void ref_moist_N2_##SFX(TF* N2, const TF* thl, const TF* dzi, TF* thvref) \
{ GEOM; calc_N2<TF>(N2, thl, dzi, thvref, RANGE); } \
MHH_EXPORT void ref_moist_buoyancy_bot_##SFX(TF* b, TF* bbot, TF* thl, TF* thlbot, TF* qt, TF* qtbot, TF* thvref, TF* thvrefh) \
{ GEOM; calc_buoyancy_bot<TF>(b, bbot, thl, thlbot, qt, qtbot, thvref, thvrefh, g.icells, g.jcells, g.icells*g.jcells, g.kstart); } \
MHH_EXPORT void ref_moist_buoyancy_fluxbot_##SFX(TF* bfluxbot, TF* thl, TF* thlfluxbot, TF* qt, TF* qtfluxbot, TF* thvrefh) \
{ GEOM; calc_buoyancy_fluxbot<TF>(bfluxbot, thl, thlfluxbot, qt, qtfluxbot, thvrefh, g.icells, g.jcells, g.kstart, g.icells*g.jcells); } \
MHH_EXPORT void ref_mean_profile_##SFX(TF* prof, const TF* fld, int itot, int jtot) \
{ GEOM; const double n = itot * jtot; \
  for (int k=0; k<g.kcells; ++k) \
  { double tmp = 0.; \
    for (int j=g.jstart; j<g.jend; ++j) for (int i=g.istart; i<g.iend; ++i) tmp += fld[i + j*g.icells + k*g.icells*g.jcells]; \
    prof[k] = tmp / n; } }

DEFINE(double, f64)
DEFINE(float, f32)
