// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// C-ABI view of the reference's tridiagonal solver `tdma` (reference src/pres_2.cxx:202-263).
// Pres_2::input/solve/output are class members that need live Grid/Fields objects; they are
// restated in oracle/oracle.py instead.
#include <src/pres_2.cxx>
#include "ref_common.h"

#define DEFINE(TF, SFX) \
MHH_EXPORT void ref_pres_2_tdma_##SFX(TF* a, TF* b, TF* c, TF* p, TF* work2d, TF* work3d, int iblock, int jblock, int kmax) \
{ tdma<TF>(a, b, c, p, work2d, work3d, iblock, jblock, kmax); }

DEFINE(double, f64)
DEFINE(float, f32)
