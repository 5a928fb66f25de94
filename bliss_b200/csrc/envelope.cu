// envelope.cu — hop energies of the envelope analyser in FP64 (kernel id BLX_K_ENVELOPE).
//
// Replaces the hot loop of reference src/tempo_atk_sort.c:109-153: normalise the whole
// interleaved int16 stream to zero mean / unit variance, and for every hop of 256 samples
//   - run the 17-tap FIR (reference include/bandpass_coeffs.h:1-7) over a 512-sample window with
//     the delay line RESTARTED at the window start,
//   - take the 512-point double real FFT,
//   - sum |X_k|^2, k = 0..256, in a FLOAT accumulator in bin order (reference :142-149).
//
// Design: for j - window_start >= 16 the restarted FIR equals the continuous
// FIR of the stream (same operands, same order), so a CTA computes the continuous FIR ONCE per
// sample for 16 consecutive hops (17 blocks of 256 samples) and only the first 16 outputs of
// each window ("heads") separately with zero history. Then 16 FFTs run at once, 16 threads each
// (fft16.cuh, double), the even/odd split writes |X_k|^2 to shared memory and one lane per hop
// replays the reference's float accumulation exactly.
//
// FP64 throughout; the summation order of the FIR is the reference's. FMA contraction is allowed
// here (the reference has none): it perturbs E[m] by ~1e-16 relative, nine orders of magnitude
// below the 1e-7 onset-count cliff measured in SURVEY.md App. B.
#include "blx_common.cuh"
#include "fft16.cuh"
#include "kernels.h"

namespace blx {

namespace {
constexpr int kEnvThreads = 256;
constexpr int kEnvH = 16;                        // hops per CTA
constexpr int kEnvSamples = (kEnvH + 1) * kHop;  // 4352 stream samples per CTA
constexpr int kPerThread = kEnvSamples / kEnvThreads; // 17 consecutive FIR outputs per thread
static_assert(kPerThread * kEnvThreads == kEnvSamples, "tile must split evenly");

constexpr int kOffC = 0;                                   // double[4352]; later P[16][257]
constexpr int kOffHeads = kOffC + kEnvSamples * 8;         // double[16][16]
constexpr int kOffXhead = kOffHeads + 16 * 16 * 8;         // double[16][16]
constexpr int kOffXchg = kOffXhead + 16 * 16 * 8;          // double2[16][272]; first the int16 tile
constexpr int kEnvSmem = kOffXchg + kEnvH * kXchgElems * 16;

// reference include/bandpass_coeffs.h:1-7 — coeffs[0][0..8]; the filter is symmetric.
__device__ __forceinline__ double fir_tap(int k) {
    constexpr double c[9] = {-0.0023470, 0.0044613, -0.0114627, 0.0226382, -0.0405147,
                             0.0580037,  -0.0779167, 0.0882711, 0.9065095};
    return c[k];
}
} // namespace

__global__ void __launch_bounds__(kEnvThreads) envelope_kernel(EnvelopeParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const SongDesc sd = p.songs[blockIdx.y];
    const int m0 = blockIdx.x * kEnvH;
    if (m0 >= sd.n_hops) return;
    const SongNorm nm = p.norm[blockIdx.y];
    if (nm.status != 0) return;
    const int h_cnt = min(kEnvH, sd.n_hops - m0);

    double *cbuf = reinterpret_cast<double *>(smem + kOffC);
    double *heads = reinterpret_cast<double *>(smem + kOffHeads);
    double *xhead = reinterpret_cast<double *>(smem + kOffXhead);
    double2 *xchg_all = reinterpret_cast<double2 *>(smem + kOffXchg);
    short *qs = reinterpret_cast<short *>(smem + kOffXchg);

    const int tid = threadIdx.x;
    const long long base = (long long)m0 * kHop; // first stream sample of this CTA
    const int n_need = (h_cnt + 1) * kHop;       // always inside the song: (m + 2) * 256 <= 512 F <= n

    // ---- phase 0: int16 tile -> shared
    if (p.dup) {
        const short *q = p.stream + sd.q_off;
        for (int i = tid; i < kEnvSamples; i += kEnvThreads) qs[i] = (i < n_need) ? q[(base + i) >> 1] : (short)0;
    } else {
        const short *q = p.stream + sd.pcm_off;
        for (int i = tid; i < kEnvSamples; i += kEnvThreads) qs[i] = (i < n_need) ? q[base + i] : (short)0;
    }
    __syncthreads();

    // ---- phase 1: normalise + continuous FIR, 17 consecutive outputs per thread
    {
        const int j0 = kPerThread * tid;
        double xv[16 + kPerThread];
#pragma unroll
        for (int i = 0; i < 16 + kPerThread; ++i) {
            const int idx = j0 - 16 + i;
            // (s / 32768 - mean_d) / var_d, reference src/tempo_atk_sort.c:110-113; the divide is
            // a multiply by the reciprocal (<= 1 ulp apart)
            const double s = int_to_double_exact((int)qs[idx < 0 ? 0 : idx]);
            const double x = fma(s, 1.0 / 32768, -nm.mean_d) * nm.inv_var_d;
            xv[i] = (idx < 0) ? 0.0 : x; // delay line starts from zero at the tile start (hop m0's window)
        }
#pragma unroll
        for (int o = 0; o < kPerThread; ++o) {
            double y = 0;
#pragma unroll
            for (int k = 7; k >= 1; --k) y += fir_tap(k) * (xv[o + 16 - k] + xv[o + k]);
            y += xv[o + 8] * fir_tap(8);
            y += fir_tap(0) * (xv[o + 16] + xv[o]);
            cbuf[j0 + o] = y;
            const int i = j0 + o;
            if ((i & (kHop - 1)) < 16 && (i >> 8) < kEnvH) xhead[(i >> 8) * 16 + (i & 15)] = xv[16 + o];
        }
    }
    __syncthreads();

    const int w = tid >> 4, lane16 = tid & 15;
    // ---- heads: first 16 outputs of windows 1..15 with zero history
    if (w >= 1) {
        const double *xh = xhead + w * 16;
        const int t = lane16;
        double y = 0;
#pragma unroll
        for (int k = 7; k >= 1; --k) {
            const double a = (t - k >= 0) ? xh[t - k] : 0.0;
            const double b = (t - 16 + k >= 0) ? xh[t - 16 + k] : 0.0;
            y += fir_tap(k) * (a + b);
        }
        y += ((t - 8 >= 0) ? xh[t - 8] : 0.0) * fir_tap(8);
        y += fir_tap(0) * (xh[t] + 0.0);
        heads[w * 16 + t] = y;
    }
    __syncthreads();

    // ---- 16 x 512-point double real FFT
    const unsigned hw_mask = 0xFFFFu << (16 * ((tid >> 4) & 1));
    const bool active = w < h_cnt;
    double2 v[16];
    if (active) {
#pragma unroll
        for (int a = 0; a < 16; ++a) v[a] = *reinterpret_cast<const double2 *>(cbuf + w * kHop + 32 * a + 2 * lane16);
        if (w >= 1 && lane16 < 8) v[0] = *reinterpret_cast<const double2 *>(heads + w * 16 + 2 * lane16);
    }
    __syncthreads(); // cbuf and the int16 tile are dead: P aliases cbuf, xchg aliases the tile

    double *P = cbuf + w * 257;
    if (active) {
        double2 *xchg = xchg_all + w * kXchgElems;
        fft256_halfwarp<double>(v, lane16, xchg, p.tw1, hw_mask);
        __syncwarp(hw_mask);
#pragma unroll
        for (int r = 0; r < 16; ++r) xchg[lane16 + 16 * fft16_out_index(r)] = v[r];
        __syncwarp(hw_mask);
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const int k = lane16 + 16 * d;
            const double2 A = v[fft16_reg_of(d)];
            if (k == 0) {
                const double x0 = A.x + A.y, xn = A.x - A.y; // X_0 and X_256 are real
                P[0] = x0 * x0;
                P[256] = xn * xn;
            } else {
                const double2 B = xchg[256 - k];
                const double2 wk = p.tw2[k];
                const double sr = A.x + B.x, si = A.y - B.y;
                const double dr = A.x - B.x, di = A.y + B.y;
                const double tr = dr * wk.x - di * wk.y;
                const double ti = dr * wk.y + di * wk.x;
                const double ar = sr + ti, ai = si - tr;
                const double br = sr - ti, bi = si + tr;
                P[k] = 0.25 * (ar * ar + ai * ai);
                P[256 - k] = 0.25 * (br * br + bi * bi);
            }
        }
        if (lane16 == 0) {
            const double2 A = v[fft16_reg_of(8)];
            P[128] = A.x * A.x + A.y * A.y;
        }
    }
    __syncthreads();

    // ---- float accumulation in bin order, one lane per hop (reference src/tempo_atk_sort.c:142-150)
    if (tid < h_cnt) {
        const double *Pw = cbuf + tid * 257;
        float sum_fft = 0.0f;
#pragma unroll 4
        for (int k = 0; k <= 256; ++k) sum_fft = (float)((double)sum_fft + Pw[k]);
        p.energy[sd.env_off + m0 + tid] = (double)sum_fft;
    }
}

cudaError_t launch_envelope(const EnvelopeParams &p, int max_hops, int n_songs, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(envelope_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kEnvSmem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (max_hops <= 0) return cudaSuccess;
    dim3 grid((unsigned)((max_hops + kEnvH - 1) / kEnvH), (unsigned)n_songs);
    envelope_kernel<<<grid, kEnvThreads, kEnvSmem, st>>>(p);
    return cudaGetLastError();
}

} // namespace blx
