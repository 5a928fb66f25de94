/* TEST INFRASTRUCTURE ONLY — CPU restatement of the decode-stage resampler (include/blx_resample.h).
 *
 * The reference calls libswresample (third-party, absent from /root/reference; unpinned apt package) at
 * reference src/decode.c:313-345 (set-up: in layout/rate/format of the file -> stereo / 22 050 Hz / s16) and
 * :388-392 (swr_convert per decoded frame, then flush). This file follows the algorithm spelled out in
 * include/blx_resample.h, one output frame at a time, in plain C. Pinned against libswresample 6.1.100 itself by
 * tools/make_golden_resample.py (build container) -> tests/golden/resample_*.npz, and against the reference's
 * md5 pins of the two resampled fixtures (reference tests/test_decode.c:35-36,55-56) in tests/test_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/blx_resample.h"

static inline int16_t clip16(long v) { return (int16_t)(v > 32767 ? 32767 : v < -32768 ? -32768 : v); }
static inline int16_t float_to_s16(float y) { return clip16(lrintf(y * 32768.0f)); }

/* one sample of the reader's array as the float libswresample computes in */
static inline float to_internal(int32_t raw, int kind, int bits, int mono) {
    float v;
    if (kind == BLX_RS_KIND_F32) memcpy(&v, &raw, 4);
    else if (kind == BLX_RS_KIND_S32) v = (float)(int32_t)((uint32_t)raw << (32 - bits)) * (1.0f / 2147483648.0f);
    else v = (float)(int16_t)((uint32_t)raw << (16 - bits)) * (1.0f / 32768.0f);
    if (mono) v *= (float)M_SQRT1_2;
    return v;
}

/* Returns the number of output frames; out (interleaved stereo, 2 * frames values) may be NULL to count. */
long long orc_resample_to_s16(const int32_t *samples, int kind, int bits, int channels, long long n_frames, int in_rate,
                              int16_t *out) {
    if (channels != 1 && channels != 2) return -1;
    const int mono = channels == 1;
    if (in_rate == BLX_RS_OUT_RATE) { /* format conversion / up-mix only */
        if (!out) return n_frames;
        for (long long t = 0; t < n_frames; ++t)
            for (int c = 0; c < 2; ++c) {
                const int32_t raw = samples[t * channels + (mono ? 0 : c)];
                int16_t v;
                if (kind == BLX_RS_KIND_U8) {
                    const int s16 = raw * 256;
                    v = mono ? clip16((s16 * 23170 + 16384) >> 15) : (int16_t)s16;
                } else if (kind == BLX_RS_KIND_S32 && !mono) {
                    v = (int16_t)((int32_t)((uint32_t)raw << (32 - bits)) >> 16);
                } else if (kind == BLX_RS_KIND_S16 && !mono) {
                    v = (int16_t)((uint32_t)raw << (16 - bits));
                } else if (kind == BLX_RS_KIND_S16) { /* int16 mono: rematrix in int16 */
                    v = clip16(((int16_t)((uint32_t)raw << (16 - bits)) * 23170 + 16384) >> 15);
                } else {
                    v = float_to_s16(to_internal(raw, kind, bits, mono));
                }
                out[2 * t + c] = v;
            }
        return n_frames;
    }
    blx_rs_plan p;
    if (blx_rs_plan_make(in_rate, BLX_RS_OUT_RATE, &p)) return -1;
    long long refl;
    const long long n_out = blx_rs_out_frames(&p, n_frames, &refl);
    if (!out) return n_out;
    const int L = p.L, L8 = (L + 7) & ~7;
    const int gain_last = mono && BLX_RS_MONO_GAIN_LAST(in_rate), gain_first = mono && !gain_last;
    if (kind == BLX_RS_KIND_U8) {
        int16_t *bank = (int16_t *)malloc(sizeof(int16_t) * (size_t)p.P * L);
        blx_rs_build_s16(&p, bank);
        for (long long m = 0; m < n_out; ++m) {
            const long long start = blx_rs_window_start(&p, m);
            const int16_t *h = bank + (size_t)blx_rs_window_phase(&p, m) * L;
            for (int c = 0; c < channels; ++c) {
                int32_t val = 1 << 14;
                for (int i = 0; i < L; ++i) {
                    int s16 = samples[blx_rs_reflect(start + i, n_frames) * channels + c] * 256;
                    if (gain_first) s16 = clip16((s16 * 23170 + 16384) >> 15);
                    val += s16 * (int32_t)h[i];
                }
                int16_t v = clip16(val >> 15);
                if (gain_last) v = clip16((v * 23170 + 16384) >> 15);
                if (mono) out[2 * m] = out[2 * m + 1] = v;
                else out[2 * m + c] = v;
            }
        }
        free(bank);
        return n_out;
    }
    float *bank = (float *)malloc(sizeof(float) * (size_t)p.P * L);
    blx_rs_build_f32(&p, bank);
    for (long long m = 0; m < n_out; ++m) {
        const long long start = blx_rs_window_start(&p, m);
        const float *h = bank + (size_t)blx_rs_window_phase(&p, m) * L;
        for (int c = 0; c < channels; ++c) {
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = 0; i < L8; ++i) {
                if (i >= L) break; /* the kernel's zero padding adds nothing */
                const float x = to_internal(samples[blx_rs_reflect(start + i, n_frames) * channels + c], kind, bits, gain_first);
                acc[i & 7] = fmaf(x, h[i], acc[i & 7]);
            }
            const float t0 = acc[0] + acc[4], t1 = acc[1] + acc[5], t2 = acc[2] + acc[6], t3 = acc[3] + acc[7];
            const float u0 = t0 + t2, u1 = t1 + t3;
            float y = u0 + u1;
            if (gain_last) y *= (float)M_SQRT1_2;
            const int16_t v = float_to_s16(y);
            if (mono) out[2 * m] = out[2 * m + 1] = v;
            else out[2 * m + c] = v;
        }
    }
    free(bank);
    return n_out;
}
