// TEST INFRASTRUCTURE -- compiles the product's __host__ __device__ moist arithmetic (microhh_b200/csrc/thermo_moist_kernels.cuh)
// for the CPU so that tests/test_moist_hostcheck.py can hold it against the oracle without a GPU.  Built on demand with nvcc.
#include "../../microhh_b200/csrc/thermo_moist_kernels.cuh"

#define EXPORT extern "C" __attribute__((visibility("default")))
using namespace mhh;

#define DEFINE(TF, SFX) \
EXPORT int hc_sat_adjust_##SFX(long n, const TF* thl, const TF* qt, TF p, TF exn, TF* ql, TF* qi, TF* t, TF* qs) \
{ int bad = 0; \
  for (long i = 0; i < n; ++i) { const SatAdjust<TF> a = moist_sat_adjust<TF>(thl[i], qt[i], p, exn); \
      ql[i] = a.ql; qi[i] = a.qi; t[i] = a.t; qs[i] = a.qs; bad += a.converged ? 0 : 1; } \
  return bad; } \
EXPORT TF hc_exner_##SFX(TF p) { return moist_exner<TF>(p); } \
EXPORT void hc_buoyancy_##SFX(long n, TF exn, const TF* thl, const TF* qt, const TF* ql, const TF* qi, TF thvref, TF* b) \
{ for (long i = 0; i < n; ++i) b[i] = moist_buoyancy<TF>(exn, thl[i], qt[i], ql[i], qi[i], thvref); } \
EXPORT void hc_buoyancy_no_ql_##SFX(long n, const TF* thl, const TF* qt, TF thvref, TF* b) \
{ for (long i = 0; i < n; ++i) b[i] = moist_buoyancy_no_ql<TF>(thl[i], qt[i], thvref); } \
EXPORT void hc_buoyancy_flux_no_ql_##SFX(long n, const TF* thl, const TF* thlflux, const TF* qt, const TF* qtflux, TF thvref, TF* b) \
{ for (long i = 0; i < n; ++i) b[i] = moist_buoyancy_flux_no_ql<TF>(thl[i], thlflux[i], qt[i], qtflux[i], thvref); } \
EXPORT int hc_base_state_##SFX(TF* pref, TF* prefh, TF* rho, TF* rhoh, TF* thv, TF* thvh, TF* ex, TF* exh, \
        const TF* thlmean, const TF* qtmean, TF pbot, int kstart, int kend, const TF* z, const TF* dz, const TF* dzh) \
{ const MoistProfiles<TF> b{pref, prefh, rho, rhoh, thv, thvh, ex, exh}; \
  return moist_base_state_serial<TF>(b, thlmean, qtmean, pbot, kstart, kend, z, dz, dzh); } \
/* host mirror of moist_base_state_kernel's skeleton (parallel sweep -> running products -> until no bit changes) */ \
EXPORT int hc_base_state_fp_##SFX(TF* pref, TF* prefh, TF* rho, TF* rhoh, TF* thv, TF* thvh, TF* ex, TF* exh, \
        const TF* thlmean, const TF* qtmean, TF pbot, int kstart, int kend, const TF* z, const TF* dz, const TF* dzh, \
        TF* F, TF* Fh, int cold, int* sweeps) \
{ const MoistProfiles<TF> b{pref, prefh, rho, rhoh, thv, thvh, ex, exh}; \
  const bool scratch = cold != 0 || !(b.pref[kstart] > TF(0.)) || !(b.prefh[kend] > TF(0.)); \
  if (scratch) for (int k = kstart; k <= kend; ++k) { b.prefh[k] = pbot; b.pref[k] = pbot; } \
  int bad = 0; *sweeps = 0; \
  for (int it = 0; it < kend - kstart + 3; ++it) \
  { bad = 0; \
    for (int k = kstart; k <= kend; ++k) bad += moist_base_level<TF>(b, thlmean, qtmean, k, kstart, kend, z, dz, dzh, F, Fh); \
    const int changed = moist_base_cumprod<TF>(b, pbot, kstart, kend, F, Fh); \
    ++*sweeps; \
    if (!changed) break; } \
  b.pref[kstart - 1] = TF(2.) * b.prefh[kstart] - b.pref[kstart]; \
  return bad; }

DEFINE(double, f64)
DEFINE(float, f32)
