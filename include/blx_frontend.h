/* BLX front-end specification v1 — OUR addition (the reference has no such stage).
 *
 * The reference's analysers consume int16 / 22 050 Hz / 2-channel interleaved PCM
 * (reference src/decode.c:7-9,187-193). BASELINE.json's workloads are 44.1 kHz mono
 * float32, which FFmpeg's libswresample would convert inside bl_audio_decode
 * (reference src/decode.c:317-345). libswresample is third-party and absent here, and
 * its SIMD polyphase filter cannot be matched bit for bit, so the engine defines its
 * own conversion and parity is stated at the analysers' input:
 *
 *   out frame t (t = 0 .. n_in/2 - 1), x[i] = 0 outside [0, n_in):
 *     a_k = x[2t - (2k+1)] + x[2t + (2k+1)]            k = 0..5   (float add)
 *     acc = C * x[2t]                                             (float mul)
 *     acc = fmaf(H_k, a_k, acc)                        k = 0..5 in order (single rounding each)
 *     q   = clamp(rintf(acc), -32768, 32767)                      (round half to even)
 *     S[2t] = S[2t+1] = (int16) q                                 (mono -> L = R)
 *
 * i.e. a 23-tap Kaiser(4.5) half-band low-pass, 2:1 decimation, with the int16
 * scale 32768 and libswresample's default -3 dB mono->stereo gain (0.70710678)
 * folded into the taps. Every operation is an IEEE-754 binary32 operation with one
 * rounding, so a C host (fmaf) and the GPU (fma.rn.f32) produce identical int16.
 * duration (whole seconds) = n_in / 44100.
 */
#ifndef BLX_FRONTEND_H
#define BLX_FRONTEND_H

#define BLX_FE_HALO 11        /* input samples needed on each side of 2t */
#define BLX_FE_NPAIRS 6
#define BLX_FE_IN_RATE 44100
#define BLX_FE_CENTER 0x1.6a09e6p+13f /* 11585.2373046875 = 0.5 * 32768 * 0.70710678 */
#define BLX_FE_TAPS                                                                          \
    {                                                                                        \
        0x1.c58670p+12f,  /* +-1 :  7256.40234375      */                                    \
        -0x1.089bacp+11f, /* +-3 : -2116.86474609375   */                                    \
        0x1.e04f60p+9f,   /* +-5 :   960.6201171875    */                                    \
        -0x1.b224b6p+8f,  /* +-7 :  -434.1434020996094 */                                    \
        0x1.49ede2p+7f,   /* +-9 :   164.96461486816406*/                                    \
        -0x1.32e188p+5f   /* +-11:   -38.36012268066406*/                                    \
    }
#endif
