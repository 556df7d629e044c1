/* Minimal single-precision FFTW3 API stand-in (FFTW 3.3.10 is the reference's
 * pinned third-party FFT, reference CMakeLists.txt:109-110; it is not
 * installed in this image).  Only the five entry points the reference calls
 * are provided (reference src/acquire.c:316-319,194,254,375-378).
 * Implementation: fftshim.c (radix-4 Stockham, float32).  Oracle build only. */
#pragma once
#include <complex.h>
#include <stddef.h>
typedef float complex fftwf_complex;
typedef struct fftshim_plan_s *fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)
fftwf_complex *fftwf_alloc_complex(size_t n);
void fftwf_free(void *p);
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
