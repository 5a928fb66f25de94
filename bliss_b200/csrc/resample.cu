// resample.cu — the decode-stage resampler of bl_audio_decode on the device (include/blx_resample.h).
//
// Replaces the libswresample calls of reference src/decode.c:313-345,388-392: any decoded file that is not
// already int16 / 22 050 Hz becomes int16 / 22 050 Hz / stereo. One thread computes one output frame of one
// channel: it gathers its window of L input frames (mirrored at both ends of the file), converts each to the
// float libswresample computes in, and sums taps * samples in the order of that library's x86 FMA3 kernel
// (eight accumulators over taps i mod 8, fused multiply-adds, fixed tree), so the int16 output is the library's
// bit for bit. The filter bank (P phases x L taps, built on the host in double) is read through L1/L2.
// Compiled with -fmad=false: every fused operation here is an explicit fmaf.
#include "blx_common.cuh"
#include "kernels.h"
#include "../../include/blx_resample.h"

namespace blx {

namespace {
__device__ __forceinline__ short clip16(int v) { return (short)max(-32768, min(32767, v)); }
__device__ __forceinline__ short float_to_s16(float y) { return clip16(__float2int_rn(__fmul_rn(y, 32768.0f))); }

__device__ __forceinline__ float to_internal(int raw, int kind, int bits, bool mono) {
    float v;
    if (kind == BLX_RS_KIND_F32) v = __int_as_float(raw);
    else if (kind == BLX_RS_KIND_S32) v = __fmul_rn(__int2float_rn((int)((unsigned)raw << (32 - bits))), 1.0f / 2147483648.0f);
    else v = __fmul_rn(__int2float_rn((int)(short)((unsigned)raw << (16 - bits))), 1.0f / 32768.0f);
    if (mono) v = __fmul_rn(v, 0.70710678118654752440f);
    return v;
}

__device__ __forceinline__ int load_raw(const ResampleParams &p, long long i) { return p.in16 ? (int)p.in16[i] : p.in[i]; }

__global__ void resample_kernel(ResampleParams p) {
    const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (m >= p.n_out) return;
    const bool mono = p.channels == 1;
    short v;
    if (p.P == 0) { // same rate: format conversion / up-mix only
        const int raw = load_raw(p, m * p.channels + (mono ? 0 : c));
        if (p.kind == BLX_RS_KIND_U8) {
            const int s16 = raw * 256;
            v = mono ? clip16((s16 * 23170 + 16384) >> 15) : (short)s16;
        } else if (p.kind == BLX_RS_KIND_S32 && !mono) {
            v = (short)((int)((unsigned)raw << (32 - p.bits)) >> 16);
        } else if (p.kind == BLX_RS_KIND_S16 && !mono) {
            v = (short)((unsigned)raw << (16 - p.bits));
        } else if (p.kind == BLX_RS_KIND_S16) {
            v = clip16(((int)(short)((unsigned)raw << (16 - p.bits)) * 23170 + 16384) >> 15);
        } else {
            v = float_to_s16(to_internal(raw, p.kind, p.bits, mono));
        }
    } else {
        const long long start = (m * p.q) / p.P - p.center;
        const int ph = (int)((m * p.q) % p.P);
        const bool gain_last = mono && p.mono_gain_last, gain_first = mono && !p.mono_gain_last;
        if (p.kind == BLX_RS_KIND_U8) {
            const short *h = p.bank_s16 + (size_t)ph * p.L;
            int val = 1 << 14;
            for (int i = 0; i < p.L; ++i) {
                const long long t = start + i;
                const long long src = t < 0 ? -t : (t < p.n_in ? t : 2 * p.n_in - 1 - t);
                int s16 = load_raw(p, src * p.channels + c) * 256;
                if (gain_first) s16 = clip16((s16 * 23170 + 16384) >> 15);
                val += s16 * (int)h[i];
            }
            v = clip16(val >> 15);
            if (gain_last) v = clip16(((int)v * 23170 + 16384) >> 15);
        } else {
            const float *h = p.bank_f32 + (size_t)ph * p.L;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int i0 = 0; i0 < p.L; i0 += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int i = i0 + j;
                    if (i < p.L) {
                        const long long t = start + i;
                        const long long src = t < 0 ? -t : (t < p.n_in ? t : 2 * p.n_in - 1 - t);
                        acc[j] = __fmaf_rn(to_internal(load_raw(p, src * p.channels + c), p.kind, p.bits, gain_first), h[i], acc[j]);
                    }
                }
            }
            const float t0 = __fadd_rn(acc[0], acc[4]), t1 = __fadd_rn(acc[1], acc[5]);
            const float t2 = __fadd_rn(acc[2], acc[6]), t3 = __fadd_rn(acc[3], acc[7]);
            float y = __fadd_rn(__fadd_rn(t0, t2), __fadd_rn(t1, t3));
            if (gain_last) y = __fmul_rn(y, 0.70710678118654752440f);
            v = float_to_s16(y);
        }
    }
    if (mono) { p.out[2 * m] = v; p.out[2 * m + 1] = v; }
    else p.out[2 * m + c] = v;
}
} // namespace

cudaError_t launch_resample(const ResampleParams &p, cudaStream_t st) {
    if (p.n_out <= 0) return cudaSuccess;
    const int threads = 256;
    dim3 grid((unsigned)((p.n_out + threads - 1) / threads), (unsigned)p.channels);
    resample_kernel<<<grid, threads, 0, st>>>(p);
    return cudaGetLastError();
}

} // namespace blx
