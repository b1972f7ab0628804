/* Stand-in for fftw3.h (FFTW itself is not in this image): lets pres_2.cxx / pres_4.cxx / fft.cxx compile.
 * Plans record their parameters; fftw_execute hands the transform to a callback registered by the test
 * (oracle/ref/ref_fftw_shim.cpp), which runs the oracle's own 1-D R2HC / HC2R -- so the reference's FFT and pressure
 * GLUE (slice loops, strides, matrix build, solver, ghost cells) runs as compiled from /root/reference while the
 * 1-D transform itself stays the restated one (see oracle/README.md). */
#ifndef MHH_ORACLE_FFTW_SHIM_H
#define MHH_ORACLE_FFTW_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
struct mhh_shim_plan_s { void* in; void* out; int n, howmany, stride, dist, kind, is_float; };
typedef struct mhh_shim_plan_s* fftw_plan;
/* a distinct type for the single-precision plan, as in FFTW */
struct mhh_shim_planf_s { struct mhh_shim_plan_s p; };
typedef struct mhh_shim_planf_s* fftwf_plan;
typedef int fftw_r2r_kind;
typedef int fftwf_r2r_kind;
#define FFTW_R2HC 0
#define FFTW_HC2R 1
#define FFTW_ESTIMATE (1U << 6)

double* fftw_alloc_real(size_t n);
float*  fftwf_alloc_real(size_t n);
void fftw_free(void* p);
void fftwf_free(void* p);
fftw_plan  fftw_plan_many_r2r(int rank, const int* n, int howmany, double* in, const int* inembed, int istride, int idist,
                              double* out, const int* onembed, int ostride, int odist, const fftw_r2r_kind* kind, unsigned flags);
fftwf_plan fftwf_plan_many_r2r(int rank, const int* n, int howmany, float* in, const int* inembed, int istride, int idist,
                               float* out, const int* onembed, int ostride, int odist, const fftwf_r2r_kind* kind, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftwf_execute(const fftwf_plan p);
void fftw_destroy_plan(fftw_plan p);
void fftwf_destroy_plan(fftwf_plan p);
void fftw_cleanup(void);
void fftwf_cleanup(void);
int  fftw_import_wisdom_from_filename(const char* filename);
int  fftwf_import_wisdom_from_filename(const char* filename);
int  fftw_export_wisdom_to_filename(const char* filename);
int  fftwf_export_wisdom_to_filename(const char* filename);
void fftw_forget_wisdom(void);
void fftwf_forget_wisdom(void);
#ifdef __cplusplus
}
#endif
#endif
