/* blx_resample.h — specification of the decode-stage resampler of bl_audio_decode.
 *
 * The reference hands every decoded file that is not already int16 / 22 050 Hz to FFmpeg's
 * libswresample with default options and asks for int16 / 22 050 Hz / stereo (reference
 * src/decode.c:313-345, 388-392); only then do the analysers see it. libswresample is a third-party
 * dependency that is absent from the reference tree (apt `libswresample-dev`, unpinned), so this file
 * restates its published algorithm (libswresample/resample.c, resample_template.c, rematrix.c,
 * audioconvert.c; defaults of options.c) as a specification of our own, shared by
 *   - the product (host: plan + filter bank, bliss_b200/host/decode.c; device: csrc/resample.cu),
 *   - the CPU checker (oracle/resample.c),
 * and validated bit for bit against libswresample 6.1.100 (the copy opencv-python-headless vendors in the
 * build container; tools/make_golden_resample.py, fixtures under tests/golden/). With it
 * bl_analyze("song_s32.flac") reproduces the reference's md5 pins of the resampled PCM (reference
 * tests/test_decode.c:35-36,55-56) and its force vector (reference tests/test_analyze.c:63-78).
 *
 * Algorithm (out_rate = 22 050, in_rate = the file's):
 *   factor = min(out_rate * 0.97 / in_rate, 1)                       cutoff 0.97
 *   L      = ceil(32 / factor), rounded up to even                    taps per phase (filter_size 32)
 *   P / q  = out_rate / in_rate in lowest terms                       P phases ("exact rational", P <= 1024)
 *   h[ph][i] = sinc(x) * I0(9 sqrt(1 - w^2)) / norm,  x = pi (i - c - ph / P) factor,  w = 2 x / (factor L pi),
 *              c = (L - 1) / 2, norm = sum_i of the un-normalised phase-0 taps       Kaiser window, beta 9
 *   output m reads input frames s .. s + L - 1, s = floor(m q / P) - c, with phase (m q) mod P;
 *   in front of the file x[-t] = x[t]; behind it x[N + j] = x[N - 1 - j] for j < r = (min(N - s*, L) + 1) / 2,
 *   s* = the first window start that does not fit in N frames; outputs run while s + L <= N + r.
 * Sample formats (what FFmpeg's FLAC / PCM decoders hand over): <= 16 bit -> int16 (left-justified),
 * 17..32 bit -> int32 (left-justified), float32. Internal arithmetic is float32 for all of them:
 *   int16 * 2^-15, int32 -> float (RN) * 2^-31, mono sources times (float)sqrt(1/2) (centre -> L, R) - before the
 *   filter, or behind it when in_rate < 11 025 Hz (libswresample resamples first when that is the cheaper order);
 *   y = sum h x in the order of libswresample's x86 FMA3 kernel: eight accumulators a[j] over taps i = j mod 8
 *   with fused multiply-adds, then ((a0 + a4) + (a2 + a6)) + ((a1 + a5) + (a3 + a7));
 *   out = clip_int16(lrintf(y * 32768)), the mono result written to both channels.
 * 8-bit sources run in int16 instead (taps lrintf(h * 32768), sum + 2^14 >> 15): same plan, BLX_RS_KIND_U8.
 * A file at 22 050 Hz that only needs a format conversion: int32 stereo -> s >> 16; float -> the last line;
 * mono -> through float as above.
 */
#ifndef BLX_RESAMPLE_H_
#define BLX_RESAMPLE_H_

#include <math.h>
#include <stdint.h>

#define BLX_RS_OUT_RATE 22050
#define BLX_RS_MAX_PHASES 1024
#define BLX_RS_MAX_TAPS 1024

/* how the reader's int32 sample array is to be read */
#define BLX_RS_KIND_S16 0 /* integer, <= 16 significant bits */
#define BLX_RS_KIND_S32 1 /* integer, 17..32 significant bits */
#define BLX_RS_KIND_F32 2 /* the bits of an IEEE float */
#define BLX_RS_KIND_U8 3  /* 8-bit PCM, already re-centred (value - 128) */

/* mono sources: 1 = the -3 dB up-mix gain is applied to the filter's output, 0 = to its input */
#define BLX_RS_MONO_GAIN_LAST(in_rate) ((in_rate) < 11025)

typedef struct blx_rs_plan {
    int in_rate, out_rate;
    int L;      /* taps per phase */
    int P;      /* phases */
    int q;      /* input frames advance by q / P per output frame */
    int center; /* (L - 1) / 2 */
    double factor;
} blx_rs_plan;

static inline long long blx_rs_gcd(long long a, long long b) {
    while (b) { const long long t = a % b; a = b; b = t; }
    return a;
}

/* 0 on success; -1 if the ratio needs more than BLX_RS_MAX_PHASES phases or L exceeds BLX_RS_MAX_TAPS */
static inline int blx_rs_plan_make(int in_rate, int out_rate, blx_rs_plan *p) {
    if (in_rate <= 0 || out_rate <= 0) return -1;
    double factor = out_rate * 0.97 / in_rate;
    if (factor > 1.0) factor = 1.0;
    int L = (int)ceil(32 / factor);
    if (L < 1) L = 1;
    if (L != 1) L = (L + 1) & ~1;
    const long long g = blx_rs_gcd(in_rate, out_rate);
    if (out_rate / g > BLX_RS_MAX_PHASES || L > BLX_RS_MAX_TAPS) return -1;
    p->in_rate = in_rate; p->out_rate = out_rate;
    p->L = L; p->P = (int)(out_rate / g); p->q = (int)(in_rate / g);
    p->center = (L - 1) / 2; p->factor = factor;
    return 0;
}

/* I0(x), power series until it stops changing */
static inline double blx_rs_bessel_i0(double x) {
    double v = 1, lastv = 0, t = 1;
    x = x * x / 4;
    for (int i = 1; v != lastv; i++) { lastv = v; t *= x / ((double)i * i); v += t; }
    return v;
}

/* Un-normalised taps of phase `ph` into tab[0..L); returns their sum. */
static inline double blx_rs_phase_taps(const blx_rs_plan *p, int ph, double *tab) {
    const double factor = p->factor;
    double s = (factor == 1.0) ? sin(M_PI * ph / p->P) * ((p->center & 1) ? 1 : -1) : 0, sum = 0;
    for (int i = 0; i < p->L; i++) {
        const double x = M_PI * ((double)(i - p->center) - (double)ph / p->P) * factor;
        double y;
        if (x == 0) y = 1.0;
        else if (factor == 1.0) y = s / x;
        else y = sin(x) / x;
        const double w = 2.0 * x / (factor * p->L * M_PI);
        y *= blx_rs_bessel_i0(9.0 * sqrt(fmax(1 - w * w, 0)));
        tab[i] = y;
        s = -s;
        sum += y;
    }
    return sum;
}

/* Filter banks, P rows of L taps: the lower half of the phases is computed, the upper half mirrored
 * (h[P - ph][L - 1 - i] = h[ph][i]) when P is even, as libswresample builds them. */
static inline void blx_rs_build_f32(const blx_rs_plan *p, float *bank) {
    double tab[BLX_RS_MAX_TAPS];
    const int ph_nb = (p->P % 2) ? p->P : p->P / 2 + 1;
    double norm = 0;
    for (int ph = 0; ph < ph_nb; ph++) {
        const double sum = blx_rs_phase_taps(p, ph, tab);
        if (!ph) norm = sum;
        for (int i = 0; i < p->L; i++) bank[(size_t)ph * p->L + i] = (float)(tab[i] * 1 / norm);
        if (p->P % 2 || ph == 0 || ph == p->P - ph) continue;
        for (int i = 0; i < p->L; i++) bank[(size_t)(p->P - ph) * p->L + p->L - 1 - i] = bank[(size_t)ph * p->L + i];
    }
}
static inline void blx_rs_build_s16(const blx_rs_plan *p, int16_t *bank) {
    double tab[BLX_RS_MAX_TAPS];
    const int ph_nb = (p->P % 2) ? p->P : p->P / 2 + 1;
    double norm = 0;
    for (int ph = 0; ph < ph_nb; ph++) {
        const double sum = blx_rs_phase_taps(p, ph, tab);
        if (!ph) norm = sum;
        for (int i = 0; i < p->L; i++) {
            const long v = lrintf((float)(tab[i] * 32768 / norm));
            bank[(size_t)ph * p->L + i] = (int16_t)(v > 32767 ? 32767 : v < -32768 ? -32768 : v);
        }
        if (p->P % 2 || ph == 0 || ph == p->P - ph) continue;
        for (int i = 0; i < p->L; i++) bank[(size_t)(p->P - ph) * p->L + p->L - 1 - i] = bank[(size_t)ph * p->L + i];
    }
}

/* First input frame of output frame m's window (may be negative) and its phase. */
static inline long long blx_rs_window_start(const blx_rs_plan *p, long long m) { return (m * p->q) / p->P - p->center; }
static inline int blx_rs_window_phase(const blx_rs_plan *p, long long m) { return (int)((m * p->q) % p->P); }

/* Output frames for N input frames, and the number of frames mirrored behind the file. */
static inline long long blx_rs_out_frames(const blx_rs_plan *p, long long N, long long *reflection) {
    if (N <= p->L) { if (reflection) *reflection = 0; return 0; } /* the library waits for L + 1 frames before its first output */
    /* first m whose window does not fit in N frames: floor(m q / P) >= N - L + center + 1 */
    const long long thr = N - p->L + p->center + 1;
    long long m = (thr * p->P + p->q - 1) / p->q;
    if (m < 0) m = 0;
    while (m > 0 && ((m - 1) * p->q) / p->P >= thr) --m;
    while ((m * p->q) / p->P < thr) ++m;
    long long leftover = N - blx_rs_window_start(p, m);
    if (leftover > p->L) leftover = p->L;
    if (leftover < 0) leftover = 0;
    const long long r = (leftover + 1) / 2, total = N + r;
    while (blx_rs_window_start(p, m) + p->L <= total) ++m;
    if (reflection) *reflection = r;
    return m;
}

/* Input frame that the window position t (start + i) reads: mirrored at both ends. */
static inline long long blx_rs_reflect(long long t, long long N) { return t < 0 ? -t : (t < N ? t : 2 * N - 1 - t); }

#endif /* BLX_RESAMPLE_H_ */
