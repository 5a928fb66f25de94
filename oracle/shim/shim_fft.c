/* TEST INFRASTRUCTURE ONLY (oracle/): never linked into the product library.
 *
 * Our own double-precision real FFT, serving the two third-party FFT APIs the
 * reference's hot path calls, so that the reference's analyser sources can be
 * compiled verbatim in a container that has neither FFmpeg nor fftw3:
 *   - libavcodec av_rdft_{init,calc,end}  (reference src/frequency_sort.c:65,83,134)
 *   - fftw_plan_dft_r2c_1d / fftw_execute (reference src/tempo_atk_sort.c:94,141)
 * Both are mathematically a plain 512-point DFT; only last-ulp behaviour is
 * library specific. This file is NOT FFmpeg or fftw3 code.
 *
 * Algorithm: N-point real FFT = N/2-point complex Stockham autosort FFT
 * (radix-4 passes, one radix-2 pass if log2(N/2) is odd) + the usual
 * even/odd split post-processing. Twiddles are precomputed per plan.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <libavformat/avformat.h>
#include <libavcodec/avfft.h>
#include <fftw3.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#if defined(__x86_64__) && defined(__GNUC__) && !defined(ORACLE_NO_CLONES)
#define SHIM_CLONES __attribute__((target_clones("avx2,fma", "default")))
#else
#define SHIM_CLONES
#endif

typedef struct {
    int n;        /* real length */
    int h;        /* n / 2 = complex length */
    double *twr;  /* cos(2 pi k / h), k < h   */
    double *twi;  /* -sin(2 pi k / h)          */
    double *ptr;  /* post-processing twiddle cos(2 pi k / n), k <= h/2 ... */
    double *pti;  /* -sin(2 pi k / n) */
    double *ar, *ai, *br, *bi; /* ping-pong work arrays, length h */
} rfft_plan;

static rfft_plan *rfft_plan_new(int n) {
    rfft_plan *p = (rfft_plan *)calloc(1, sizeof(*p));
    p->n = n;
    p->h = n / 2;
    p->twr = (double *)malloc(sizeof(double) * p->h);
    p->twi = (double *)malloc(sizeof(double) * p->h);
    p->ptr = (double *)malloc(sizeof(double) * (p->h + 1));
    p->pti = (double *)malloc(sizeof(double) * (p->h + 1));
    p->ar = (double *)malloc(sizeof(double) * p->h);
    p->ai = (double *)malloc(sizeof(double) * p->h);
    p->br = (double *)malloc(sizeof(double) * p->h);
    p->bi = (double *)malloc(sizeof(double) * p->h);
    for (int k = 0; k < p->h; ++k) {
        p->twr[k] = cos(2.0 * M_PI * k / p->h);
        p->twi[k] = -sin(2.0 * M_PI * k / p->h);
    }
    for (int k = 0; k <= p->h; ++k) {
        p->ptr[k] = cos(2.0 * M_PI * k / n);
        p->pti[k] = -sin(2.0 * M_PI * k / n);
    }
    return p;
}

static void rfft_plan_free(rfft_plan *p) {
    if (!p) return;
    free(p->twr); free(p->twi); free(p->ptr); free(p->pti);
    free(p->ar); free(p->ai); free(p->br); free(p->bi);
    free(p);
}

/* Complex forward FFT of length h held in (ar, ai); result left in (ar, ai). */
SHIM_CLONES
static void cfft_forward(rfft_plan *p) {
    const int h = p->h;
    double *xr = p->ar, *xi = p->ai, *yr = p->br, *yi = p->bi;
    int l = h, m = 1;
    while (l >= 4) {
        l /= 4;
        const int tstep = h / (4 * l); /* twiddle index step: w = exp(-2 pi i j /(4 l)) */
        for (int j = 0; j < l; ++j) {
            const double w1r = p->twr[j * tstep], w1i = p->twi[j * tstep];
            const double w2r = p->twr[2 * j * tstep], w2i = p->twi[2 * j * tstep];
            const double w3r = p->twr[3 * j * tstep], w3i = p->twi[3 * j * tstep];
            const double *x0r = xr + m * j, *x0i = xi + m * j;
            const double *x1r = x0r + m * l, *x1i = x0i + m * l;
            const double *x2r = x1r + m * l, *x2i = x1i + m * l;
            const double *x3r = x2r + m * l, *x3i = x2i + m * l;
            double *y0r = yr + m * 4 * j, *y0i = yi + m * 4 * j;
            double *y1r = y0r + m, *y1i = y0i + m;
            double *y2r = y1r + m, *y2i = y1i + m;
            double *y3r = y2r + m, *y3i = y2i + m;
            for (int k = 0; k < m; ++k) {
                const double d0r = x0r[k] + x2r[k], d0i = x0i[k] + x2i[k];
                const double d1r = x0r[k] - x2r[k], d1i = x0i[k] - x2i[k];
                const double d2r = x1r[k] + x3r[k], d2i = x1i[k] + x3i[k];
                /* d3 = -i (x1 - x3) */
                const double d3r = x1i[k] - x3i[k], d3i = -(x1r[k] - x3r[k]);
                const double e1r = d1r + d3r, e1i = d1i + d3i;
                const double e2r = d0r - d2r, e2i = d0i - d2i;
                const double e3r = d1r - d3r, e3i = d1i - d3i;
                y0r[k] = d0r + d2r;
                y0i[k] = d0i + d2i;
                y1r[k] = e1r * w1r - e1i * w1i;
                y1i[k] = e1r * w1i + e1i * w1r;
                y2r[k] = e2r * w2r - e2i * w2i;
                y2i[k] = e2r * w2i + e2i * w2r;
                y3r[k] = e3r * w3r - e3i * w3i;
                y3i[k] = e3r * w3i + e3i * w3r;
            }
        }
        m *= 4;
        double *t;
        t = xr; xr = yr; yr = t;
        t = xi; xi = yi; yi = t;
    }
    if (l == 2) { /* final radix-2 pass, twiddle = 1 */
        for (int k = 0; k < m; ++k) {
            const double a_r = xr[k], a_i = xi[k], b_r = xr[k + m], b_i = xi[k + m];
            yr[k] = a_r + b_r; yi[k] = a_i + b_i;
            yr[k + m] = a_r - b_r; yi[k + m] = a_i - b_i;
        }
        double *t;
        t = xr; xr = yr; yr = t;
        t = xi; xi = yi; yi = t;
    }
    if (xr != p->ar) {
        memcpy(p->ar, xr, sizeof(double) * h);
        memcpy(p->ai, xi, sizeof(double) * h);
    }
}

/* Real forward FFT: in[n] -> (re[k], im[k]), k = 0..n/2. X_k = sum x_t e^{-2 pi i k t / n}. */
SHIM_CLONES
static void rfft_forward(rfft_plan *p, const double *in, double *re, double *im) {
    const int h = p->h;
    for (int t = 0; t < h; ++t) {
        p->ar[t] = in[2 * t];
        p->ai[t] = in[2 * t + 1];
    }
    cfft_forward(p);
    const double *zr = p->ar, *zi = p->ai;
    re[0] = zr[0] + zi[0];
    im[0] = 0.0;
    re[h] = zr[0] - zi[0];
    im[h] = 0.0;
    for (int k = 1; k < h; ++k) {
        const double ar_ = zr[k], ai_ = zi[k];
        const double br_ = zr[h - k], bi_ = -zi[h - k]; /* conj(Z[h-k]) */
        const double er = 0.5 * (ar_ + br_), ei = 0.5 * (ai_ + bi_);   /* even part */
        const double dr = 0.5 * (ar_ - br_), di = 0.5 * (ai_ - bi_);   /* (Z_k - conj Z_{h-k}) / 2 */
        /* odd part = -i * w^k * d, w^k = ptr + i pti */
        const double tr = dr * p->ptr[k] - di * p->pti[k];
        const double ti = dr * p->pti[k] + di * p->ptr[k];
        re[k] = er + ti;
        im[k] = ei - tr;
    }
}

/* ------------------------------------------------------------------ */
/* libavutil / libavcodec entry points                                 */
/* ------------------------------------------------------------------ */
void *av_malloc(size_t size) { return malloc(size); }
void av_free(void *ptr) { free(ptr); }

struct RDFTContext {
    rfft_plan *plan;
    double *in, *re, *im;
};

RDFTContext *av_rdft_init(int nbits, enum RDFTransformType trans) {
    if (trans != DFT_R2C) return NULL;
    RDFTContext *s = (RDFTContext *)calloc(1, sizeof(*s));
    const int n = 1 << nbits;
    s->plan = rfft_plan_new(n);
    s->in = (double *)malloc(sizeof(double) * n);
    s->re = (double *)malloc(sizeof(double) * (n / 2 + 1));
    s->im = (double *)malloc(sizeof(double) * (n / 2 + 1));
    return s;
}

/* In-place, packed output as the reference expects (comment at
 * src/frequency_sort.c:87): data[0] = Re X_0, data[1] = Re X_{n/2},
 * data[2k], data[2k+1] = Re X_k, Im X_k. Computed in double, rounded to float. */
void av_rdft_calc(RDFTContext *s, FFTSample *data) {
    const int n = s->plan->n, h = n / 2;
    for (int t = 0; t < n; ++t) s->in[t] = (double)data[t];
    rfft_forward(s->plan, s->in, s->re, s->im);
    data[0] = (float)s->re[0];
    data[1] = (float)s->re[h];
    for (int k = 1; k < h; ++k) {
        data[2 * k] = (float)s->re[k];
        data[2 * k + 1] = (float)s->im[k];
    }
}

void av_rdft_end(RDFTContext *s) {
    if (!s) return;
    rfft_plan_free(s->plan);
    free(s->in); free(s->re); free(s->im);
    free(s);
}

/* ------------------------------------------------------------------ */
/* fftw3 entry points                                                  */
/* ------------------------------------------------------------------ */
struct shim_fftw_plan_s {
    rfft_plan *plan;
    double *in;
    fftw_complex *out;
    double *re, *im;
};

void *fftw_malloc(size_t n) { return malloc(n); }
void fftw_free(void *p) { free(p); }

fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags) {
    (void)flags;
    fftw_plan p = (fftw_plan)calloc(1, sizeof(*p));
    p->plan = rfft_plan_new(n);
    p->in = in;
    p->out = out;
    p->re = (double *)malloc(sizeof(double) * (n / 2 + 1));
    p->im = (double *)malloc(sizeof(double) * (n / 2 + 1));
    return p;
}

void fftw_execute(const fftw_plan p) {
    const int h = p->plan->n / 2;
    rfft_forward(p->plan, p->in, p->re, p->im);
    for (int k = 0; k <= h; ++k) {
        p->out[k][0] = p->re[k];
        p->out[k][1] = p->im[k];
    }
}

void fftw_destroy_plan(fftw_plan p) {
    if (!p) return;
    rfft_plan_free(p->plan);
    free(p->re); free(p->im);
    free(p);
}

void fftw_cleanup(void) {}

/* Exposed for the oracle restatement (oracle/bliss_oracle.c) and for tests. */
void *orc_rfft_new(int n) { return rfft_plan_new(n); }
void orc_rfft_free(void *p) { rfft_plan_free((rfft_plan *)p); }
void orc_rfft_exec(void *p, const double *in, double *re, double *im) {
    rfft_forward((rfft_plan *)p, in, re, im);
}
