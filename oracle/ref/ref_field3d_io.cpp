// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's restart IO of one 3-D field, Field3d_io<TF>::save_field3d / load_field3d (serial build,
// src/field3d_io.cxx:669-751): a REAL Field3d_io<TF> object (its constructor only stores the two references,
// src/field3d_io.cxx:33-37) on the zeroed Master image and the reference's own Grid<TF> of ref_grid.cpp.
#include <src/field3d_io.cxx>
#include "ref_common.h"

#define DEFINE(TF, SFX) \
MHH_EXPORT int ref_field3d_save_##SFX(TF* data, TF* tmp1, TF* tmp2, const char* filename, TF offset, int kstart, int kend) \
{ Field3d_io<TF> io(*static_cast<Master*>(ref_master_image()), *static_cast<Grid<TF>*>(ref_grid_image(sizeof(TF) == 4))); \
  return io.save_field3d(data, tmp1, tmp2, filename, offset, kstart, kend); } \
MHH_EXPORT int ref_field3d_load_##SFX(TF* data, TF* tmp1, TF* tmp2, const char* filename, TF offset, int kstart, int kend) \
{ Field3d_io<TF> io(*static_cast<Master*>(ref_master_image()), *static_cast<Grid<TF>*>(ref_grid_image(sizeof(TF) == 4))); \
  return io.load_field3d(data, tmp1, tmp2, filename, offset, kstart, kend); }

DEFINE(double, f64)
DEFINE(float, f32)
