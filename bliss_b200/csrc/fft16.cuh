// fft16.cuh — register-resident 16-point complex FFT and the 256-point / real-512
// building blocks shared by the float (frequency) and double (envelope) kernels.
//
// A 512-point real FFT is computed as a 256-point complex FFT of z[n] = x[2n] + i x[2n+1]
// followed by the even/odd split. The 256-point FFT is 16 x 16: sixteen threads of one
// half-warp each hold 16 complex points in registers, do a 16-point FFT, exchange through
// shared memory (one transpose) and do a second 16-point FFT. All synchronisation inside
// one transform is __syncwarp().
//
// This replaces libavcodec's av_rdft (reference src/frequency_sort.c:83) and fftw3's r2c
// plan (reference src/tempo_atk_sort.c:141); both are a plain forward DFT
// X_k = sum_t x_t exp(-2 pi i k t / 512).
#pragma once
#include <cuda_runtime.h>

namespace blx {

template <typename T> struct cplx;
template <> struct cplx<float> { using type = float2; };
template <> struct cplx<double> { using type = double2; };

template <typename T> __device__ __forceinline__ typename cplx<T>::type mk(T x, T y) {
    typename cplx<T>::type r; r.x = x; r.y = y; return r;
}

// (a.x + i a.y) * (w.x + i w.y)
template <typename C> __device__ __forceinline__ C cmul(C a, C w) {
    C r;
    r.x = a.x * w.x - a.y * w.y;
    r.y = a.x * w.y + a.y * w.x;
    return r;
}

// Packed single precision (sm_100 add/sub.rn.f32x2 -> FADD2): a float2 is one 64-bit register pair, so a
// complex addition is ONE instruction. Every lane is the same IEEE addition as the scalar form.
__device__ __forceinline__ float2 cadd2(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
    return *reinterpret_cast<float2 *>(&r);
}
__device__ __forceinline__ float2 csub2(float2 a, float2 b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
    return *reinterpret_cast<float2 *>(&r);
}

// forward 4-point DFT in place (W4 = -i)
template <typename C> __device__ __forceinline__ void dft4(C &a0, C &a1, C &a2, C &a3) {
    C t0, t1, t2, t3;
    t0.x = a0.x + a2.x; t0.y = a0.y + a2.y;
    t1.x = a0.x - a2.x; t1.y = a0.y - a2.y;
    t2.x = a1.x + a3.x; t2.y = a1.y + a3.y;
    t3.x = a1.x - a3.x; t3.y = a1.y - a3.y;
    a0.x = t0.x + t2.x; a0.y = t0.y + t2.y;
    a2.x = t0.x - t2.x; a2.y = t0.y - t2.y;
    a1.x = t1.x + t3.y; a1.y = t1.y - t3.x; // t1 - i t3
    a3.x = t1.x - t3.y; a3.y = t1.y + t3.x; // t1 + i t3
}

// the same for float2 with packed additions: 6 FADD2 + 4 FADD instead of 16 FADD
template <> __device__ __forceinline__ void dft4<float2>(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    const float2 t0 = cadd2(a0, a2), t1 = csub2(a0, a2), t2 = cadd2(a1, a3), t3 = csub2(a1, a3);
    a0 = cadd2(t0, t2);
    a2 = csub2(t0, t2);
    a1.x = t1.x + t3.y; a1.y = t1.y - t3.x; // t1 - i t3
    a3.x = t1.x - t3.y; a3.y = t1.y + t3.x; // t1 + i t3
}

// Forward 16-point FFT in registers. Input v[n], n = 0..15 natural order.
// Output: register r holds X[(r >> 2) + 4 * (r & 3)]  (base-4 digit reversal).
template <typename T> __device__ __forceinline__ void fft16(typename cplx<T>::type (&v)[16]) {
    using C = typename cplx<T>::type;
    const T c1 = T(0.92387953251128675613);  // cos(pi/8)
    const T s1 = T(0.38268343236508977173);  // sin(pi/8)
    const T h = T(0.70710678118654752440);   // sqrt(1/2)
    // stage 1: for each n2, 4-point DFT over n1 of v[4 n1 + n2]; result A[n2][k1] in v[4 k1 + n2]
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dft4<C>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
    // twiddles W16^(n2 k1) on v[4 k1 + n2]
    // k1 = 1: m = n2          -> W^1, W^2, W^3
    v[5] = cmul<C>(v[5], mk<T>(c1, -s1));
    { C a = v[6]; v[6].x = (a.x + a.y) * h; v[6].y = (a.y - a.x) * h; }           // W^2 = h(1 - i)
    v[7] = cmul<C>(v[7], mk<T>(s1, -c1));
    // k1 = 2: m = 2 n2        -> W^2, W^4, W^6
    { C a = v[9]; v[9].x = (a.x + a.y) * h; v[9].y = (a.y - a.x) * h; }
    { C a = v[10]; v[10].x = a.y; v[10].y = -a.x; }                                 // W^4 = -i
    { C a = v[11]; v[11].x = (a.y - a.x) * h; v[11].y = -(a.x + a.y) * h; }         // W^6 = -h(1 + i)
    // k1 = 3: m = 3 n2        -> W^3, W^6, W^9
    v[13] = cmul<C>(v[13], mk<T>(s1, -c1));
    { C a = v[14]; v[14].x = (a.y - a.x) * h; v[14].y = -(a.x + a.y) * h; }
    v[15] = cmul<C>(v[15], mk<T>(-c1, s1));
    // stage 2: for each k1, 4-point DFT over n2 of v[4 k1 + n2]; X[k1 + 4 k2] in v[4 k1 + k2]
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4<C>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}

// index held by register r after fft16
__host__ __device__ constexpr int fft16_out_index(int r) { return (r >> 2) + 4 * (r & 3); }
// register that holds output index d after fft16
__host__ __device__ constexpr int fft16_reg_of(int d) { return 4 * (d & 3) + (d >> 2); }

// Exchange buffer geometry: 16 rows (c) of 17 complex (b, one pad) per transform.
constexpr int kXchgRow = 17;
constexpr int kXchgElems = 16 * kXchgRow; // complex elements per transform (>= 257 needed for Z)

// 256-point complex FFT across the 16 threads of a half-warp.
//   lane16: this thread's index b (0..15) inside the transform
//   v: in  v[a] = z[16 a + lane16]
//      out register r holds Z[lane16 + 16 * fft16_out_index(r)]
//   xchg: this transform's private shared buffer of kXchgElems complex
//   tw1: shared table, tw1[c * 16 + b] = exp(-2 pi i b c / 256)
//   mask: the 16 lanes taking part (0x0000FFFF or 0xFFFF0000)
// The caller must __syncwarp(mask) before reusing xchg.
template <typename T, bool TW_ON_LOAD = false>
__device__ __forceinline__ void fft256_halfwarp(typename cplx<T>::type (&v)[16], int lane16,
                                                typename cplx<T>::type *xchg,
                                                const typename cplx<T>::type *tw1, unsigned mask) {
    using C = typename cplx<T>::type;
    fft16<T>(v);
    // The twiddle W256^(b c) between the passes is applied by the lane that WRITES element (c, b) (it holds all 16
    // outputs, so the table reads compete with 64 live registers) or, TW_ON_LOAD, by the lane that READS it (the table is
    // symmetric in b and c; the same products, the registers are free at that point so all reads are in flight at once).
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int c = fft16_out_index(r);
        C u = (c == 0 || TW_ON_LOAD) ? v[r] : cmul<C>(v[r], tw1[c * 16 + lane16]);
        xchg[c * kXchgRow + lane16] = u;
    }
    __syncwarp(mask);
    if (TW_ON_LOAD) {
        C w[16];
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            v[b] = xchg[lane16 * kXchgRow + b];
            if (b > 0) w[b] = tw1[b * 16 + lane16];
        }
#pragma unroll
        for (int b = 1; b < 16; ++b) v[b] = cmul<C>(v[b], w[b]);
    } else {
#pragma unroll
        for (int b = 0; b < 16; ++b) v[b] = xchg[lane16 * kXchgRow + b];
    }
    fft16<T>(v);
}

// The same with the inter-pass twiddles of this lane already in registers: twr[c] = exp(-2 pi i lane16 c / 256), c = 1..15
// (a kernel that walks many transforms loads them once; float only - 30 registers).
template <typename T>
__device__ __forceinline__ void fft256_halfwarp_regtw(typename cplx<T>::type (&v)[16], int lane16, typename cplx<T>::type *xchg,
                                                      const typename cplx<T>::type (&twr)[16], unsigned mask) {
    using C = typename cplx<T>::type;
    fft16<T>(v);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int c = fft16_out_index(r);
        xchg[c * kXchgRow + lane16] = (c == 0) ? v[r] : cmul<C>(v[r], twr[c]);
    }
    __syncwarp(mask);
#pragma unroll
    for (int b = 0; b < 16; ++b) v[b] = xchg[lane16 * kXchgRow + b];
    fft16<T>(v);
}

} // namespace blx
