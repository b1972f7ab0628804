// TEST INFRASTRUCTURE ONLY -- single-precision (USESP) instance of the reference's FFT<TF> (see ref_fft_impl.h).
#define FLOAT_SINGLE
#include "ref_fft_impl.h"
DEFINE_FFT(float, f32)
