// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
// The reference's FFT<TF> (src/fft.cxx) compiled where it lies, over the FFTW shim (shim/fftw3.h): constructor, init(),
// load() (plan creation, src/fft.cxx:118-160 / 163-205) and the serial exec_forward / exec_backward slice loops
// (src/fft.cxx:338-452) are the reference's own; only the batched 1-D transform is the registered callback.
// src/fft.cxx instantiates ONE precision per translation unit (FLOAT_SINGLE), hence ref_fft.cpp and ref_fft_f32.cpp.
#include <src/fft.cxx>
#include "ref_common.h"

#define DEFINE_FFT(TF, SFX) \
MHH_EXPORT void* ref_fft_create_##SFX() \
{ \
    return new FFT<TF>(*static_cast<Master*>(ref_master_image()), *static_cast<Grid<TF>*>(ref_grid_image(sizeof(TF) == 4))); \
} \
MHH_EXPORT void ref_fft_init_##SFX(void* f) { static_cast<FFT<TF>*>(f)->init(); } \
MHH_EXPORT void ref_fft_load_##SFX(void* f) { static_cast<FFT<TF>*>(f)->load(); } \
MHH_EXPORT void ref_fft_forward_##SFX(void* f, TF* data, TF* tmp) { static_cast<FFT<TF>*>(f)->exec_forward(data, tmp); } \
MHH_EXPORT void ref_fft_backward_##SFX(void* f, TF* data, TF* tmp) { static_cast<FFT<TF>*>(f)->exec_backward(data, tmp); } \
MHH_EXPORT void ref_fft_destroy_##SFX(void* f) { delete static_cast<FFT<TF>*>(f); }
