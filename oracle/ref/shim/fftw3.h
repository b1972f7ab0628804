/* Type-only stand-in for fftw3.h: lets pres_2.cxx / pres_4.cxx / fft.h parse.
 * The oracle never calls FFTW; its FFT is restated separately (see oracle/README.md). */
#ifndef MHH_ORACLE_FFTW_SHIM_H
#define MHH_ORACLE_FFTW_SHIM_H
typedef struct mhh_shim_plan_s*  fftw_plan;
typedef struct mhh_shim_planf_s* fftwf_plan;
typedef int fftw_r2r_kind;
typedef int fftwf_r2r_kind;
#define FFTW_R2HC 0
#define FFTW_HC2R 1
#define FFTW_ESTIMATE (1U << 6)
#endif
