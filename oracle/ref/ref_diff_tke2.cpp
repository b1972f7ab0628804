// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's Deardorff SGS-TKE CPU kernels (src/diff_tke2.cxx:48-512): enforce_min_sgstke,
// calc_evisc_neutral, calc_evisc, calc_evisc_heat, sgstke_shear_tend, sgstke_buoy_tend, sgstke_diss_tend,
// sgstke_diss_tend_neutral; and of Limiter's tendency_limiter (src/limiter.cxx:35-59).  Diff_tke2::exec itself only calls
// Diff_kernels::diff_u / diff_v / diff_w / diff_c (ref_diff_smag2.cpp exports those) with tPr = 1 and the evisc / eviscs
// selection restated in oracle/step.py.
// Two reference translation units are included (their anonymous namespaces merge; the names do not clash).
#include <src/diff_tke2.cxx>
#include <src/limiter.cxx>
#include "ref_common.h"

#define GEOM const Ref_geom& g = ref_geom
#define RANGE g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells
#define RANGE3 g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.jcells, g.icells*g.jcells

template<typename TF> static Boundary_cyclic<TF>& cyclic_stub_tke2()
{
    alignas(16) static char buf[sizeof(Boundary_cyclic<TF>)];
    return *reinterpret_cast<Boundary_cyclic<TF>*>(buf);
}

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_tke2_enforce_min_##SFX(TF* sgstke) \
{ GEOM; enforce_min_sgstke<TF>(sgstke, RANGE3, cyclic_stub_tke2<TF>()); } \
MHH_EXPORT void ref_tke2_evisc_neutral_##SFX(TF* evisc, const TF* sgstke, const TF* u, const TF* v, const TF* w, const TF* z, const TF* dz, \
        const TF* z0m, TF dx, TF dy, TF cn, TF cm, int mason) \
{ GEOM; Boundary_cyclic<TF>& bc = cyclic_stub_tke2<TF>(); \
  if (mason) calc_evisc_neutral<TF, Surface_model::Enabled, true >(evisc, sgstke, u, v, w, z, dz, z0m, dx, dy, cn, cm, RANGE3, bc); \
  else       calc_evisc_neutral<TF, Surface_model::Enabled, false>(evisc, sgstke, u, v, w, z, dz, z0m, dx, dy, cn, cm, RANGE3, bc); } \
MHH_EXPORT void ref_tke2_evisc_##SFX(TF* evisc, const TF* sgstke, const TF* u, const TF* v, const TF* w, const TF* N2, const TF* bgradbot, \
        const TF* z, const TF* dz, const TF* z0m, TF dx, TF dy, TF cn, TF cm, int mason) \
{ GEOM; Boundary_cyclic<TF>& bc = cyclic_stub_tke2<TF>(); \
  if (mason) calc_evisc<TF, Surface_model::Enabled, true >(evisc, sgstke, u, v, w, N2, bgradbot, z, dz, z0m, dx, dy, cn, cm, RANGE3, bc); \
  else       calc_evisc<TF, Surface_model::Enabled, false>(evisc, sgstke, u, v, w, N2, bgradbot, z, dz, z0m, dx, dy, cn, cm, RANGE3, bc); } \
MHH_EXPORT void ref_tke2_evisc_heat_##SFX(TF* evisch, const TF* evisc, const TF* sgstke, const TF* N2, const TF* bgradbot, \
        const TF* z, const TF* dz, const TF* z0m, TF dx, TF dy, TF cn, TF ch1, TF ch2, int mason) \
{ GEOM; Boundary_cyclic<TF>& bc = cyclic_stub_tke2<TF>(); \
  if (mason) calc_evisc_heat<TF, Surface_model::Enabled, true >(evisch, evisc, sgstke, N2, bgradbot, z, dz, z0m, dx, dy, cn, ch1, ch2, RANGE3, bc); \
  else       calc_evisc_heat<TF, Surface_model::Enabled, false>(evisch, evisc, sgstke, N2, bgradbot, z, dz, z0m, dx, dy, cn, ch1, ch2, RANGE3, bc); } \
MHH_EXPORT void ref_tke2_shear_tend_##SFX(TF* at, const TF* a, const TF* evisc, const TF* strain2) \
{ GEOM; sgstke_shear_tend<TF>(at, a, evisc, strain2, RANGE); } \
MHH_EXPORT void ref_tke2_buoy_tend_##SFX(TF* at, const TF* a, const TF* evisch, const TF* N2, const TF* bgradbot) \
{ GEOM; sgstke_buoy_tend<TF>(at, a, evisch, N2, bgradbot, RANGE); } \
MHH_EXPORT void ref_tke2_diss_tend_##SFX(TF* at, const TF* a, const TF* N2, const TF* bgradbot, const TF* z, const TF* dz, const TF* z0m, \
        TF dx, TF dy, TF cn, TF ce1, TF ce2, int mason) \
{ GEOM; sgstke_diss_tend<TF>(at, a, N2, bgradbot, z, dz, z0m, dx, dy, cn, ce1, ce2, RANGE, mason != 0); } \
MHH_EXPORT void ref_tke2_diss_tend_neutral_##SFX(TF* at, const TF* a, const TF* z, const TF* dz, const TF* z0m, TF dx, TF dy, TF ce1, TF ce2, int mason) \
{ GEOM; sgstke_diss_tend_neutral<TF>(at, a, z, dz, z0m, dx, dy, ce1, ce2, RANGE, mason != 0); } \
MHH_EXPORT void ref_limiter_##SFX(TF* at, const TF* a, TF min_value, TF dt) \
{ GEOM; tendency_limiter<TF>(at, a, min_value, dt, RANGE); }

DEFINE(double, f64)
DEFINE(float, f32)
