// TEST INFRASTRUCTURE ONLY -- double-precision instance of the reference's FFT<TF> (see ref_fft_impl.h).
#include "ref_fft_impl.h"
DEFINE_FFT(double, f64)
