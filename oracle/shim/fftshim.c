/* Single-precision complex FFT behind the five-function FFTW3 API subset the
 * reference uses (see fftw3.h in this directory).  Algorithm: out-of-place
 * Stockham autosort, radix-4 stages with one trailing radix-2 stage when
 * log2(n) is odd, float32 twiddles computed in double.  Written for the
 * oracle build; not part of the product. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "fftw3.h"

struct fftshim_plan_s {
    int n;
    int sign;
    float complex *in, *out;
    float complex *tw;      /* tw[k] = exp(sign * 2*pi*i*k/n), k < n */
    float complex *scratch; /* n elements */
};

fftwf_complex *fftwf_alloc_complex(size_t n)
{
    void *p = NULL;
    if (posix_memalign(&p, 64, n * sizeof(fftwf_complex)) != 0)
        return NULL;
    return (fftwf_complex *)p;
}

void fftwf_free(void *p) { free(p); }

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags)
{
    (void)flags;
    if (n < 2 || (n & (n - 1)))
        return NULL;
    fftwf_plan p = (fftwf_plan)calloc(1, sizeof(*p));
    p->n = n;
    p->sign = sign < 0 ? -1 : 1;
    p->in = in;
    p->out = out;
    p->tw = fftwf_alloc_complex(n);
    p->scratch = fftwf_alloc_complex(n);
    for (int k = 0; k < n; k++) {
        double a = p->sign * 2.0 * M_PI * (double)k / (double)n;
        p->tw[k] = (float)cos(a) + (float)sin(a) * I;
    }
    return p;
}

void fftwf_destroy_plan(fftwf_plan p)
{
    if (!p) return;
    fftwf_free(p->tw);
    fftwf_free(p->scratch);
    free(p);
}

/* multiply by sign*j */
static inline float complex rot90(float complex v, int sign)
{
    return sign < 0 ? CMPLXF(cimagf(v), -crealf(v)) : CMPLXF(-cimagf(v), crealf(v));
}

static void stage4(int n, int s, int N, int sign, const float complex *tw,
                   const float complex *restrict x, float complex *restrict y)
{
    (void)N;
    const int n1 = n / 4;
    for (int p = 0; p < n1; p++) {
        const float complex w1 = tw[p * s];
        const float complex w2 = tw[2 * p * s];
        const float complex w3 = tw[3 * p * s];
        for (int q = 0; q < s; q++) {
            const float complex a = x[q + s * p];
            const float complex b = x[q + s * (p + n1)];
            const float complex c = x[q + s * (p + 2 * n1)];
            const float complex d = x[q + s * (p + 3 * n1)];
            const float complex apc = a + c, amc = a - c, bpd = b + d;
            const float complex jbmd = rot90(b - d, sign);
            y[q + s * (4 * p + 0)] = apc + bpd;
            y[q + s * (4 * p + 1)] = w1 * (amc + jbmd);
            y[q + s * (4 * p + 2)] = w2 * (apc - bpd);
            y[q + s * (4 * p + 3)] = w3 * (amc - jbmd);
        }
    }
}

static void stage2(int n, int s, const float complex *tw,
                   const float complex *restrict x, float complex *restrict y)
{
    const int m = n / 2;
    for (int p = 0; p < m; p++) {
        const float complex w = tw[p * s];
        for (int q = 0; q < s; q++) {
            const float complex a = x[q + s * p];
            const float complex b = x[q + s * (p + m)];
            y[q + s * (2 * p + 0)] = a + b;
            y[q + s * (2 * p + 1)] = (a - b) * w;
        }
    }
}

void fftwf_execute(const fftwf_plan p)
{
    const int N = p->n;
    int nstages = 0;
    for (int n = N; n > 1; n = (n >= 4) ? n / 4 : n / 2)
        nstages++;
    /* ping-pong so the last stage writes p->out */
    float complex *bufs[2];
    bufs[(nstages & 1)] = p->out;       /* stage k writes bufs[(k+1)&1 ...] below */
    bufs[!(nstages & 1)] = p->scratch;
    const float complex *src = p->in;
    int n = N, s = 1, k = 0;
    while (n > 1) {
        float complex *dst = bufs[(k + 1) & 1];
        if (n >= 4) {
            stage4(n, s, N, p->sign, p->tw, src, dst);
            n /= 4; s *= 4;
        } else {
            stage2(n, s, p->tw, src, dst);
            n /= 2; s *= 2;
        }
        src = dst;
        k++;
    }
    if (src != p->out)
        memcpy(p->out, src, sizeof(float complex) * N);
}
