/* TEST INFRASTRUCTURE ONLY — CPU restatement of the BLX front-end specification v1
 * (include/blx_frontend.h): 44.1 kHz mono float32 -> int16 / 22 050 Hz / stereo, the
 * format the reference's analysers take (reference src/decode.c:7-9,187-193). The GPU
 * kernels implement the same specification; tests check the int16 streams are equal. */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include "../include/blx_frontend.h"

static inline float fe_at(const float *x, long n, long i) { return (i >= 0 && i < n) ? x[i] : 0.0f; }

/* out must hold 2 * (n_in / 2) int16. Returns nSamples (= count of int16 written). */
int orc_frontend_f32(const float *x, long n_in, int16_t *out) {
    static const float H[BLX_FE_NPAIRS] = BLX_FE_TAPS;
    const long n_out = n_in / 2;
    for (long t = 0; t < n_out; ++t) {
        float acc = BLX_FE_CENTER * fe_at(x, n_in, 2 * t);
        for (int k = 0; k < BLX_FE_NPAIRS; ++k) {
            float a = fe_at(x, n_in, 2 * t - (2 * k + 1)) + fe_at(x, n_in, 2 * t + (2 * k + 1));
            acc = fmaf(H[k], a, acc);
        }
        float q = rintf(acc);
        if (q > 32767.0f) q = 32767.0f;
        if (q < -32768.0f) q = -32768.0f;
        out[2 * t] = out[2 * t + 1] = (int16_t)q;
    }
    return (int)(2 * n_out);
}
