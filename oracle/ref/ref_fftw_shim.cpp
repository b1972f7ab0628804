// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never of the product path.
//
// The FFTW entry points the reference's src/fft.cxx calls, for an image without FFTW: plans record the
// fftw_plan_many_r2r parameters and fftw_execute forwards them to a callback the TEST registers
// (ref_set_fft_callback), which performs the batched 1-D R2HC / HC2R with the oracle's restated transform.
// Nothing here is reference code; it only lets the reference's own FFT / Pres_2 / Pres_4 glue run unmodified.
#include <cstdlib>
#include <cstdio>
#include "fftw3.h"
#include "ref_common.h"

typedef void (*mhh_fft_cb)(void* in, void* out, int n, int howmany, int stride, int dist, int kind, int is_float);
static mhh_fft_cb fft_cb = nullptr;
MHH_EXPORT void ref_set_fft_callback(mhh_fft_cb cb) { fft_cb = cb; }

static void run(const mhh_shim_plan_s* p)
{
    if (!fft_cb) { std::fprintf(stderr, "oracle fftw shim: no FFT callback registered\n"); std::abort(); }
    fft_cb(p->in, p->out, p->n, p->howmany, p->stride, p->dist, p->kind, p->is_float);
}

extern "C" {
double* fftw_alloc_real(size_t n) { return static_cast<double*>(std::calloc(n ? n : 1, sizeof(double))); }
float*  fftwf_alloc_real(size_t n) { return static_cast<float*>(std::calloc(n ? n : 1, sizeof(float))); }
void fftw_free(void* p) { std::free(p); }
void fftwf_free(void* p) { std::free(p); }
fftw_plan fftw_plan_many_r2r(int, const int* n, int howmany, double* in, const int*, int istride, int idist,
                             double* out, const int*, int, int, const fftw_r2r_kind* kind, unsigned)
{ return new mhh_shim_plan_s{in, out, n[0], howmany, istride, idist, kind[0], 0}; }
fftwf_plan fftwf_plan_many_r2r(int, const int* n, int howmany, float* in, const int*, int istride, int idist,
                               float* out, const int*, int, int, const fftwf_r2r_kind* kind, unsigned)
{ return new mhh_shim_planf_s{{in, out, n[0], howmany, istride, idist, kind[0], 1}}; }
void fftw_execute(const fftw_plan p) { run(p); }
void fftwf_execute(const fftwf_plan p) { run(&p->p); }
void fftw_destroy_plan(fftw_plan p) { delete p; }
void fftwf_destroy_plan(fftwf_plan p) { delete p; }
void fftw_cleanup(void) {}
void fftwf_cleanup(void) {}
int  fftw_import_wisdom_from_filename(const char*) { return 1; }
int  fftwf_import_wisdom_from_filename(const char*) { return 1; }
int  fftw_export_wisdom_to_filename(const char*) { return 1; }
int  fftwf_export_wisdom_to_filename(const char*) { return 1; }
void fftw_forget_wisdom(void) {}
void fftwf_forget_wisdom(void) {}
}
