/* TEST INFRASTRUCTURE ONLY — the CPU oracle. Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this; the product
 * (bliss_b200/) never does.
 *
 * A plain-C restatement of the per-song analysis of Polochon-street/bliss
 * (reference @ 20c4536), written from the reference's arithmetic,
 * not copied from it. Unlike the reference it exposes the intermediate quantities
 * (per-bin power sums, hop energies E[m], beat count) that the CUDA kernels are
 * checked against one by one.
 *
 * Parity status: PINNED.
 *   - against the reference's own golden vectors (reference tests/test_analyze.c:30-45,
 *     63-78) via tests/test_oracle.py, and
 *   - bit-for-bit against the reference's analyser sources compiled verbatim
 *     (oracle/_ref/libbliss_ref.so, same shim FFT) on seeded random songs.
 *
 * Platform contract (that of the goldens): x86-64 / SSE2 (FLT_EVAL_METHOD 0),
 * -std=c99 => no FMA contraction; float expressions are evaluated in float, mixed
 * float/double expressions in double.
 *
 * Third-party arithmetic on the path that is NOT under /root/reference: libavcodec
 * av_rdft (float 512-pt R2C; reference src/frequency_sort.c:65,83) and fftw3
 * (double 512-pt R2C; reference src/tempo_atk_sort.c:94,141), both unpinned apt
 * packages in the reference's CI. Both are a plain DFT; shim/shim_fft.c computes it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

void *orc_rfft_new(int n);
void orc_rfft_free(void *p);
void orc_rfft_exec(void *p, const double *in, double *re, double *im);

#define ORC_WIN 512

/* ================================================================== */
/* A.1 frequency rating — reference src/frequency_sort.c:20-140        */
/* ================================================================== */

/* Hann window, symmetric (N-1) form: cosine in double, stored as float
 * (reference src/frequency_sort.c:40-42). */
void orc_hann512(float *w) {
    for (int i = 0; i < ORC_WIN; ++i)
        w[i] = (float)(0.5 * (1.0 - cos(2 * M_PI * i / (ORC_WIN - 1))));
}

/* Steps 2-5: per-bin power accumulated over all frames, float accumulator, frames
 * in order (reference src/frequency_sort.c:50,67-94). ps has 257 entries; ps[0] is
 * the last frame's X0^2 (assigned, never used), ps[256] stays 0. Returns n_frames. */
int orc_frequency_spectrum(const int16_t *S, int n, int channels, float *ps) {
    float hann[ORC_WIN];
    double in[ORC_WIN], re[ORC_WIN / 2 + 1], im[ORC_WIN / 2 + 1];
    float x[ORC_WIN];
    void *plan = orc_rfft_new(ORC_WIN);
    const int n_frames = (n / channels) / ORC_WIN;
    orc_hann512(hann);
    for (int d = 0; d <= ORC_WIN / 2; ++d) ps[d] = 0.0f;
    for (int f = 0; f < n_frames; ++f) {
        const int16_t *p = S + (size_t)f * ORC_WIN * channels;
        if (channels == 2) {
            for (int d = 0; d < ORC_WIN; ++d) {
                int m = ((int)p[2 * d] + (int)p[2 * d + 1]) / 2; /* C truncation toward zero */
                x[d] = (float)m * hann[d];
            }
        } else {
            for (int d = 0; d < ORC_WIN; ++d) x[d] = (float)p[d] * hann[d];
        }
        /* av_rdft: float in, float out. Computed in double, rounded to float. */
        for (int d = 0; d < ORC_WIN; ++d) in[d] = (double)x[d];
        orc_rfft_exec(plan, in, re, im);
        {
            float x0 = (float)re[0];
            ps[0] = x0 * x0;
        }
        for (int d = 1; d < ORC_WIN / 2; ++d) {
            float fr = (float)re[d], fi = (float)im[d];
            float raw = (fr * fr) + (fi * fi);
            ps[d] += raw;
        }
    }
    orc_rfft_free(plan);
    return n_frames;
}

/* Steps 6-9: dB relative to the peak, five band means, rating
 * (reference src/frequency_sort.c:97-139). ps[1..256] are modified in place. */
float orc_frequency_from_spectrum(float *ps) {
    float peak = 0;
    float bands[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    for (int d = 1; d <= ORC_WIN / 2; ++d) {
        ps[d] = (float)sqrt((double)(ps[d] / (float)ORC_WIN));
        peak = (float)fmax((double)ps[d], (double)peak);
    }
    for (int d = 1; d <= ORC_WIN / 2; ++d)
        ps[d] = (float)(20 * log10((double)(ps[d] / peak)) - 3);
    bands[0] = (ps[2] + ps[4]) / 2;
    bands[1] = (ps[6] + ps[8]) / 2;
    for (int i = 10; i <= 60; ++i) bands[2] += ps[i];
    bands[2] /= 50;  /* 51 terms over 50: part of the contract */
    for (int i = 61; i <= 118; ++i) bands[3] += ps[i];
    bands[3] /= 57;  /* 58 terms over 57 */
    for (int i = 119; i <= 234; ++i) bands[4] += ps[i];
    bands[4] /= 115; /* 116 terms over 115 */
    float bands_sum = bands[4] + bands[3] + bands[2] - bands[0] - bands[1];
    return (float)((1. / 3.) * (double)bands_sum + 68. / 3.);
}

float orc_frequency(const int16_t *S, int n, int channels) {
    float ps[ORC_WIN / 2 + 1];
    orc_frequency_spectrum(S, n, channels, ps);
    return orc_frequency_from_spectrum(ps);
}

/* ================================================================== */
/* A.2 amplitude rating — reference src/amplitude_sort.c:12-80         */
/* ================================================================== */
#define ORC_HIST 65536
#define ORC_PASSES 301 /* for (g = 0; g <= 300; ++g), reference src/amplitude_sort.c:41 */

/* First / last non-zero sample over ALL interleaved samples
 * (reference src/amplitude_sort.c:26-31). Undefined for an all-zero song there;
 * here returns -1. */
int orc_amplitude_bounds(const int16_t *S, int n, int *start, int *end) {
    int s = 0, e = n - 1;
    while (s < n && S[s] == 0) ++s;
    if (s == n) return -1;
    while (S[e] == 0) --e;
    *start = s;
    *end = e;
    return 0;
}

/* Smoothing + normalisation + window integral on a given raw histogram of float
 * counts (reference src/amplitude_sort.c:41-79). */
float orc_amplitude_from_histogram(const float *hist_in, int start, int end) {
    float *h = (float *)malloc(sizeof(float) * ORC_HIST);
    float *sm = (float *)calloc(ORC_HIST, sizeof(float));
    memcpy(h, hist_in, sizeof(float) * ORC_HIST);
    for (int g = 0; g < ORC_PASSES; ++g) {
        sm[0] = h[0];
        sm[1] = (float)(1. / 4. * (double)(h[0] + (2 * h[1]) + h[2]));
        sm[2] = (float)(1. / 9. * (double)(h[0] + (2 * h[1]) + (3 * h[2]) + (2 * h[3]) + h[4]));
        for (int i = 3; i < ORC_HIST - 5; ++i) {
            float taps = h[i - 3] + (3 * h[i - 2]) + (6 * h[i - 1]) + (7 * h[i]) + (6 * h[i + 1]) +
                         (3 * h[i + 2]) + h[i + 3]; /* float, left to right */
            sm[i] = (float)(1. / 27. * (double)taps);
        }
        for (int i = 3; i < ORC_HIST - 5; ++i) h[i] = sm[i];
    }
    float integral = 0;
    const float span = (float)(start - end); /* negative; NOT the sample count */
    for (int i = 32767 - 1000; i <= 32767 + 1000; ++i) {
        float v = sm[i] / span;
        v = (float)((double)v * 100.);
        v = fabsf(v);
        integral += v;
    }
    free(h);
    free(sm);
    return -0.2f * integral + 6.0f;
}

float orc_amplitude(const int16_t *S, int n) {
    int start, end;
    if (orc_amplitude_bounds(S, n, &start, &end)) return NAN;
    float *hist = (float *)calloc(ORC_HIST, sizeof(float));
    for (int i = start; i <= end; ++i) hist[(int)S[i] + 32768] += 1;
    float r = orc_amplitude_from_histogram(hist, start, end);
    free(hist);
    return r;
}

/* ================================================================== */
/* A.3 envelope (tempo + attack) — reference src/tempo_atk_sort.c       */
/* ================================================================== */

/* Filter tables: the literals of reference include/bandpass_coeffs.h:1-7,484-492
 * (data, 5-7 significant digits; must be identical, not re-derived). */
static const double orc_fir[17] = {-0.0023470, 0.0044613,  -0.0114627, 0.0226382, -0.0405147, 0.0580037,
                                   -0.0779167, 0.0882711,  0.9065095,  0.0882711, -0.0779167, 0.0580037,
                                   -0.0405147, 0.0226382,  -0.0114627, 0.0044613, -0.0023470};
static const double orc_lp_b[7] = {1.9510e-05, 1.1706e-04, 2.9266e-04, 3.9021e-04,
                                   2.9266e-04, 1.1706e-04, 1.9510e-05};
static const double orc_lp_a[7] = {1.00000, -4.59007, 8.91034, -9.34191, 5.56998, -1.78845, 0.24136};

/* reference src/helpers.c:30-37: `int` accumulator, C truncating division. The
 * reference's signed overflow (|sum| >= 2^31) is UB; we define it as the wrap gcc
 * produces in practice. */
int orc_mean(const int16_t *S, int n) {
    uint32_t acc = 0;
    for (int i = 0; i < n; ++i) acc += (uint32_t)(int32_t)S[i];
    return (int32_t)acc / n;
}

/* reference src/helpers.c:39-49: int32 deviations, int64 accumulator. */
int orc_variance(const int16_t *S, int n, int mean) {
    int64_t acc = 0;
    for (int i = 0; i < n; ++i) {
        int64_t v = (int64_t)S[i] - mean;
        acc += v * v;
    }
    return (int)(acc / n);
}

/* Steps 1-5: normalise, per-hop restarted 17-tap FIR, 512-pt double FFT, power
 * summed in a FLOAT accumulator in bin order (reference src/tempo_atk_sort.c:101-153).
 * E must hold nb_frames = 2*floor(n/512) doubles; the last two stay 0. */
int orc_envelope_energy(const int16_t *S, int n, double *E) {
    const int F = n / ORC_WIN;
    const int nb_frames = 2 * F;
    const int hops = 2 * F - 2;
    const int mean = orc_mean(S, n);
    const int var = orc_variance(S, n, mean);
    const double mean_d = (double)mean / 32768;
    double var_d = (double)var / 32768;
    var_d /= 32768;
    double in[ORC_WIN], re[ORC_WIN / 2 + 1], im[ORC_WIN / 2 + 1];
    double hist[ORC_WIN + 16];
    void *plan = orc_rfft_new(ORC_WIN);
    for (int m = 0; m < nb_frames; ++m) E[m] = 0.0;
    for (int m = 0; m < hops; ++m) {
        const int16_t *p = S + (size_t)m * (ORC_WIN / 2);
        /* x[t] for t < 0 (before the window) is zero: the delay line restarts. */
        for (int t = 0; t < 16; ++t) hist[t] = 0.0;
        for (int t = 0; t < ORC_WIN; ++t) hist[16 + t] = ((double)p[t] / 32768 - mean_d) / var_d;
        for (int t = 0; t < ORC_WIN; ++t) {
            const double *x = hist + 16 + t; /* x[0] = newest, x[-k] = k samples ago */
            double y = 0;
            for (int k = 7; k >= 1; --k) y += orc_fir[k] * (x[-k] + x[-16 + k]);
            y += x[-8] * orc_fir[8];
            y += orc_fir[0] * (x[0] + x[-16]);
            in[t] = y;
        }
        orc_rfft_exec(plan, in, re, im);
        float sum_fft = 0;
        for (int k = 0; k <= ORC_WIN / 2; ++k) {
            double pw = re[k] * re[k] + im[k] * im[k];
            sum_fft = (float)((double)sum_fft + pw);
        }
        E[m] = (double)sum_fft;
    }
    orc_rfft_free(plan);
    return nb_frames;
}

/* reference src/tempo_atk_sort.c:19-40. */
void orc_rectangular_filter(double *out, const double *in, int n, int width) {
    const int half = (int)round(width / 2.);
    double run = 0;
    for (int k = 0; k < width; ++k) run += in[k];
    for (int k = 0; k < n - width; ++k) {
        out[k + half - 1] = run;
        run -= in[k];
        run += in[k + width];
    }
    for (int k = n - width; k < n; ++k) out[n - half] += in[k];
    for (int k = 0; k < n; ++k) out[k] /= width;
}

/* Steps 6-13 (reference src/tempo_atk_sort.c:170-287). If ss_out != NULL it receives
 * the final smoothed signal (n2 doubles). */
void orc_envelope_tail(const double *E, int nb_frames, int n_samples, uint64_t duration, int *beat_out,
                       double *atk_sum_out, float *tempo, float *attack, double *ss_out) {
    const int n2 = 2 * nb_frames;
    double *t1 = (double *)calloc((size_t)n2, sizeof(double));
    double *t2 = (double *)calloc((size_t)n2, sizeof(double));
    double *wa = (double *)calloc((size_t)n2, sizeof(double));
    double *ss = (double *)calloc((size_t)n2, sizeof(double));
    const float mu = 100.0f;
    const float lambda = 0.8f;
    for (int j = 0; j < nb_frames; ++j) {
        t1[2 * j] = log(1 + (double)mu * E[j]) / log((double)(1 + mu));
        t1[2 * j + 1] = 0;
    }
    /* 6th-order IIR low-pass, direct form I, zero state */
    double xh[7] = {0, 0, 0, 0, 0, 0, 0}, yh[7] = {0, 0, 0, 0, 0, 0, 0};
    double y = 0;
    for (int j = 0; j < n2; ++j) {
        for (int k = 6; k >= 1; --k) { xh[k] = xh[k - 1]; yh[k] = yh[k - 1]; }
        xh[0] = t1[j];
        yh[0] = y; /* yh[k-1] = y[j-k] */
        double d = 0, c = 0;
        for (int k = 0; k < 7; ++k) d += orc_lp_b[k] * xh[k];
        for (int k = 1; k < 7; ++k) c += orc_lp_a[k] * yh[k - 1];
        y = (d - c) / orc_lp_a[0];
        t2[j] = y;
    }
    /* half-wave rectified difference */
    t1[0] = t2[0];
    for (int j = 1; j < n2; ++j) {
        double df = t2[j] - t2[j - 1];
        t1[j] = (df > 0) ? df : 0;
    }
    const double w_lp = (double)(1 - lambda);       /* float expression 1 - 0.8f */
    const double w_df = (double)(lambda * 172);     /* float expression 0.8f * 172 */
    for (int j = 0; j < n2; ++j) wa[j] = w_lp * t2[j] + w_df * t1[j] / 10;
    double atk_sum = 0;
    for (int j = 0; j < n2 - 1; ++j) atk_sum += wa[j];
    for (int j = 0; j < n2 - 1; ++j) ss[j] += wa[j];
    orc_rectangular_filter(wa, ss, n2, 19);
    for (int k = 0; k < n2; ++k) ss[k] = 0;
    orc_rectangular_filter(ss, wa, n2, 19);
    const float epsilon = 0.000001f;
    int beat = 0;
    for (int j = 1; j < n2 - 1; ++j)
        if (((ss[j] - ss[j - 1]) > epsilon) && ((ss[j] - ss[j + 1]) > epsilon)) beat++;
    double tempo_score = (double)(4 * (float)beat / (float)duration) - 30.4;
    double atk_score = -1.74 * atk_sum * 10000 / n_samples + 58.3;
    if (beat_out) *beat_out = beat;
    if (atk_sum_out) *atk_sum_out = atk_sum;
    if (tempo) *tempo = (float)tempo_score;
    if (attack) *attack = (float)atk_score;
    if (ss_out) memcpy(ss_out, ss, sizeof(double) * (size_t)n2);
    free(t1); free(t2); free(wa); free(ss);
}

void orc_envelope(const int16_t *S, int n, uint64_t duration, float *tempo, float *attack, int *beat) {
    const int nb_frames = 2 * (n / ORC_WIN);
    double *E = (double *)malloc(sizeof(double) * (size_t)(nb_frames > 0 ? nb_frames : 1));
    orc_envelope_energy(S, n, E);
    orc_envelope_tail(E, nb_frames, n, duration, beat, NULL, tempo, attack, NULL);
    free(E);
}

/* ================================================================== */
/* A.4 rating and distances — reference src/analyze.c:63-79,88-103,127-145 */
/* ================================================================== */
typedef struct orc_result {
    float force;
    float tempo, amplitude, frequency, attack;
    int calm_or_loud; /* BL_LOUD 0 / BL_CALM 1 / BL_UNKNOWN 2 */
    int beat;
} orc_result;

float orc_rating(float tempo, float amplitude, float frequency, float attack, int *calm_or_loud) {
    float rating = (float)(fmax((double)tempo, 0) + (double)amplitude + (double)frequency + fmax((double)attack, 0));
    if (calm_or_loud) *calm_or_loud = (rating > 0) ? 0 : (rating < 0) ? 1 : 2;
    return rating;
}

void orc_analyze(const int16_t *S, int n, int channels, uint64_t duration, orc_result *r) {
    r->amplitude = orc_amplitude(S, n);
    r->frequency = orc_frequency(S, n, channels);
    orc_envelope(S, n, duration, &r->tempo, &r->attack, &r->beat);
    r->force = orc_rating(r->tempo, r->amplitude, r->frequency, r->attack, &r->calm_or_loud);
}

/* Differences, squares and the three adds in float, left to right; sqrt in double,
 * rounded to float (reference src/analyze.c:96-100). */
float orc_distance(const float *a, const float *b) {
    float s = (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]) +
              (a[3] - b[3]) * (a[3] - b[3]);
    return (float)sqrt((double)s);
}

/* reference src/analyze.c:135-142: float dot and norms, double sqrt and divide. */
float orc_cosine_similarity(const float *a, const float *b) {
    float dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
    float na = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3];
    float nb = b[0] * b[0] + b[1] * b[1] + b[2] * b[2] + b[3] * b[3];
    return (float)((double)dot / (sqrt((double)na) * sqrt((double)nb)));
}

void orc_distance_matrix(const float *v, int n, float *out) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) out[(size_t)i * n + j] = orc_distance(v + 4 * i, v + 4 * j);
}
