// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's vertical ghost-cell kernels: calc_ghost_cells_bot_2nd / _top_2nd
// (reference src/boundary.cxx:700-772), calc_ghost_cells_bot_4th / _top_4th (:776-848) and the no-penetration
// ghost cells of w, conservation and normal type (:850-922).
// bc: 0 = Dirichlet, 1 = Neumann/Flux.
#include <src/boundary.cxx>
#include "ref_common.h"

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_ghost_cells_bot_2nd_##SFX(TF* a, const TF* dzh, int bc, TF* abot, TF* agradbot) \
{ const Ref_geom& g = ref_geom; \
  calc_ghost_cells_bot_2nd<TF>(a, dzh, bc == 0 ? Boundary_type::Dirichlet_type : Boundary_type::Neumann_type, abot, agradbot, \
          g.kstart, g.icells, g.jcells, g.icells*g.jcells); } \
MHH_EXPORT void ref_ghost_cells_top_2nd_##SFX(TF* a, const TF* dzh, int bc, TF* atop, TF* agradtop) \
{ const Ref_geom& g = ref_geom; \
  calc_ghost_cells_top_2nd<TF>(a, dzh, bc == 0 ? Boundary_type::Dirichlet_type : Boundary_type::Neumann_type, atop, agradtop, \
          g.kend, g.icells, g.jcells, g.icells*g.jcells); } \
MHH_EXPORT void ref_ghost_cells_bot_4th_##SFX(TF* a, const TF* z, int bc, TF* abot, TF* agradbot) \
{ const Ref_geom& g = ref_geom; \
  calc_ghost_cells_bot_4th<TF>(a, z, bc == 0 ? Boundary_type::Dirichlet_type : Boundary_type::Neumann_type, abot, agradbot, \
          g.kstart, g.icells, g.jcells, g.icells*g.jcells); } \
MHH_EXPORT void ref_ghost_cells_top_4th_##SFX(TF* a, const TF* z, int bc, TF* atop, TF* agradtop) \
{ const Ref_geom& g = ref_geom; \
  calc_ghost_cells_top_4th<TF>(a, z, bc == 0 ? Boundary_type::Dirichlet_type : Boundary_type::Neumann_type, atop, agradtop, \
          g.kend, g.icells, g.jcells, g.icells*g.jcells); } \
MHH_EXPORT void ref_ghost_cells_w_4th_##SFX(TF* w, int conservation) \
{ const Ref_geom& g = ref_geom; const int ij = g.icells*g.jcells; \
  if (conservation) { calc_ghost_cells_botw_cons_4th<TF>(w, g.kstart, g.icells, g.jcells, ij); calc_ghost_cells_topw_cons_4th<TF>(w, g.kend, g.icells, g.jcells, ij); } \
  else { calc_ghost_cells_botw_4th<TF>(w, g.kstart, g.icells, g.jcells, ij); calc_ghost_cells_topw_4th<TF>(w, g.kend, g.icells, g.jcells, ij); } }

DEFINE(double, f64)
DEFINE(float, f32)
