// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's damping-layer and large-scale forcing kernels:
//   calc_buffer (src/buffer.cxx:38-58); add_pressure_force / enforce_fixed_flux (src/force.cxx:47-75), calc_coriolis_2nd
//   (:78-108), calc_coriolis_4th (:110-152), calc_large_scale_source (:154-170), advec_wls_2nd_local (:238-272).
// Two reference translation units are included (their anonymous namespaces merge; the names do not clash).
#include <src/buffer.cxx>
#include <src/force.cxx>
#include "ref_common.h"

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_buffer_##SFX(TF* at, const TF* a, const TF* abuf, const TF* z, TF zstart, TF zsize, TF beta, TF sigma, int bufferkstart) \
{ const Ref_geom& g = ref_geom; \
  calc_buffer<TF>(at, a, abuf, z, zstart, zsize, beta, sigma, g.istart, g.iend, g.icells, g.jstart, g.jend, g.icells*g.jcells, bufferkstart, g.kend); } \
MHH_EXPORT void ref_force_fixed_flux_##SFX(TF* ut, TF u_flux, TF u_mean, TF ut_mean, TF u_grid, TF dt) \
{ const Ref_geom& g = ref_geom; \
  enforce_fixed_flux<TF>(ut, u_flux, u_mean, ut_mean, u_grid, dt, g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells); } \
MHH_EXPORT void ref_force_pressure_##SFX(TF* ut, TF fbody) \
{ const Ref_geom& g = ref_geom; \
  add_pressure_force<TF>(ut, fbody, g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells); } \
MHH_EXPORT void ref_force_coriolis_##SFX(TF* ut, TF* vt, const TF* u, const TF* v, const TF* ug, const TF* vg, TF fc, TF ugrid, TF vgrid, int order) \
{ const Ref_geom& g = ref_geom; \
  if (order == 4) calc_coriolis_4th<TF>(ut, vt, u, v, ug, vg, fc, ugrid, vgrid, g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells); \
  else calc_coriolis_2nd<TF>(ut, vt, u, v, ug, vg, fc, ugrid, vgrid, g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells); } \
MHH_EXPORT void ref_force_ls_source_##SFX(TF* st, const TF* sls) \
{ const Ref_geom& g = ref_geom; \
  calc_large_scale_source<TF>(st, sls, g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells); } \
MHH_EXPORT void ref_force_wls_local_##SFX(TF* st, const TF* s, const TF* wls, const TF* dzhi) \
{ const Ref_geom& g = ref_geom; \
  advec_wls_2nd_local<TF>(st, s, wls, dzhi, g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells); }

DEFINE(double, f64)
DEFINE(float, f32)
