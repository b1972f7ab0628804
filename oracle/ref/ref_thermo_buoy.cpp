// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's Thermo_buoy CPU kernels (src/thermo_buoy.cxx:41-296): calc_N2, calc_buoyancy_tend_2nd / _4th,
// the slope variants calc_buoyancy_tend_u / _w / _b (2nd and 4th order) and calc_baroclinic_2nd / _4th.
#include <src/thermo_buoy.cxx>
#include "ref_common.h"

#define GEOM const Ref_geom& g = ref_geom
#define RANGE g.istart, g.iend, g.jstart, g.jend, g.kstart, g.kend, g.icells, g.icells*g.jcells

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_thermo_buoy_N2_##SFX(TF* N2, const TF* b, TF bg_n2, const TF* dzi) \
{ GEOM; calc_N2<TF>(N2, b, bg_n2, dzi, RANGE, g.kcells); } \
MHH_EXPORT void ref_thermo_buoy_tend_##SFX(TF* wt, TF* b, int order) \
{ GEOM; if (order == 4) calc_buoyancy_tend_4th<TF>(wt, b, RANGE); else calc_buoyancy_tend_2nd<TF>(wt, b, RANGE); } \
MHH_EXPORT void ref_thermo_buoy_tend_slope_##SFX(TF* ut, TF* wt, TF* bt, TF* b, TF* u, TF* w, TF alpha, TF n2, TF utrans, int order) \
{ GEOM; \
  if (order == 4) { calc_buoyancy_tend_u_4th<TF>(ut, b, alpha, RANGE); calc_buoyancy_tend_w_4th<TF>(wt, b, alpha, RANGE); \
                    calc_buoyancy_tend_b_4th<TF>(bt, u, w, alpha, n2, utrans, RANGE); } \
  else            { calc_buoyancy_tend_u_2nd<TF>(ut, b, alpha, RANGE); calc_buoyancy_tend_w_2nd<TF>(wt, b, alpha, RANGE); \
                    calc_buoyancy_tend_b_2nd<TF>(bt, u, w, alpha, n2, utrans, RANGE); } } \
MHH_EXPORT void ref_thermo_buoy_baroclinic_##SFX(TF* bt, const TF* v, TF dbdy_ls, int order) \
{ GEOM; if (order == 4) calc_baroclinic_4th<TF>(bt, v, dbdy_ls, RANGE); else calc_baroclinic_2nd<TF>(bt, v, dbdy_ls, RANGE); }

DEFINE(double, f64)
DEFINE(float, f32)
