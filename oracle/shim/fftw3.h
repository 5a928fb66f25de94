/* TEST INFRASTRUCTURE ONLY (oracle/). Declarations of the fftw3 subset used at
 * src/tempo_atk_sort.c:61,75,86-87,94,141,242-243,290-295 of the reference.
 * The implementation (shim_fft.c) is our own FFT, NOT fftw3. */
#ifndef ORACLE_SHIM_FFTW3_H
#define ORACLE_SHIM_FFTW3_H
#include <stddef.h>
typedef double fftw_complex[2];
typedef struct shim_fftw_plan_s *fftw_plan;
#define FFTW_ESTIMATE (1U << 6)
void *fftw_malloc(size_t n);
void fftw_free(void *p);
fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
void fftw_cleanup(void);
#endif
