// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's 2nd-order vertical ghost-cell kernels
// calc_ghost_cells_bot_2nd / calc_ghost_cells_top_2nd (reference src/boundary.cxx:700-772).
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
          g.kend, g.icells, g.jcells, g.icells*g.jcells); }

DEFINE(double, f64)
DEFINE(float, f32)
