/* TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
 * Element-wise std::pow through the C library, so that the numpy restatement evaluates
 * `std::pow(TF, TF)` (reference src/diff_smag2.cxx:212, 232-233, 245, 258-259) with the same
 * libm the reference's CPU build links, instead of numpy's own SIMD pow (1 ulp apart now and then). */
#include <math.h>
#define EXPORT __attribute__((visibility("default")))
EXPORT void vpow_f64(const double* x, double y, double* out, long n) { for (long i = 0; i < n; ++i) out[i] = pow(x[i], y); }
EXPORT void vpow_f32(const float* x, float y, float* out, long n) { for (long i = 0; i < n; ++i) out[i] = powf(x[i], y); }
/* std::exp for calc_evisc_neutral's van Driest damping (reference src/diff_smag2.cxx:99-100) */
EXPORT void vexp_f64(const double* x, double* out, long n) { for (long i = 0; i < n; ++i) out[i] = exp(x[i]); }
EXPORT void vexp_f32(const float* x, float* out, long n) { for (long i = 0; i < n; ++i) out[i] = expf(x[i]); }
