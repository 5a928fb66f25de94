// envelope.cu — hop energies of the envelope analyser in FP64 (kernel id BLX_K_ENVELOPE).
//
// Replaces the hot loop of reference src/tempo_atk_sort.c:109-153: normalise the whole
// interleaved int16 stream to zero mean / unit variance, and for every hop of 256 samples
//   - run the 17-tap FIR (reference include/bandpass_coeffs.h:1-7) over a 512-sample window with
//     the delay line RESTARTED at the window start,
//   - take the 512-point double real FFT,
//   - sum |X_k|^2, k = 0..256, in a FLOAT accumulator in bin order (reference :142-149).
//
// Design. For j - window_start >= 16 the restarted FIR equals the continuous
// FIR of the stream (same operands, same order), so a tile of 16 consecutive hops computes the
// continuous FIR ONCE per sample (17 blocks of 256 samples) and only the first 16 outputs of each
// window ("heads") separately with zero history. Then 16 FFTs run at once, 16 threads each
// (fft16.cuh, double). A CTA walks kTilesPerCta tiles; the int16 tile of t+1 is fetched by one 1-D
// bulk copy (cp.async.bulk / UBLKCP) as soon as the FIR of tile t has consumed the staging buffer.
//
// The float accumulation s <- (float)((double)s + p_k), k = 0..256, is a chain of 257 dependent
// roundings (~50-75 cycles each on the FP64 + conversion pipes): run naively it idles the SM. It is
// evaluated here by the 16 lanes that hold the hop's spectrum, as a SCAN:
//   while s stays inside one binade [2^e, 2^(e+1)) its float grid is g = 2^(e-23) and
//   RN_g(s + p) = s + RN_g(p) because s is a multiple of g; so with q = s / g (a 24-bit integer)
//   the chain is q += I_k with I_k = RN_g(p_k) / g, which one double addition p_k + 1.5 * 2^52 * g
//   leaves in the low mantissa word. Each lane converts its 16 bins, an integer prefix scan over the
//   half-warp gives every partial sum at once, and the first bin k* at which q + prefix reaches 2^24
//   (the sum leaves the binade) is found with a ballot. Bins below k* are final; bin k* is added with
//   the reference's own double-add / float-convert sequence, which yields the next binade, and the
//   scan restarts after k*. A hop needs one round per binade crossing (typically 2-6).
// The result equals the reference's chain except when s + p_k lies within 2^-29 grid units of a
// rounding midpoint (the reference rounds to double, then to float; exact ties follow the parity of
// the magic constant instead of q).
//
// FP64 throughout; the summation order of the FIR is the reference's. FMA contraction is allowed
// here (the reference has none): it perturbs E[m] by ~1e-16 relative, nine orders of magnitude
// below the 1e-7 onset-count cliff measured in SURVEY.md App. B.
#include "blx_common.cuh"
#include "fft16.cuh"
#include "kernels.h"

namespace blx {

namespace {
constexpr int kEnvThreads = 256;                 // 8 warps = 16 half-warps = 16 FFTs
constexpr int kEnvH = 16;                        // hops per tile
constexpr int kTilesPerCta = 8;
constexpr int kEnvSamples = (kEnvH + 1) * kHop;  // 4352 stream samples per tile
constexpr int kPerThread = kEnvSamples / kEnvThreads; // 17 consecutive FIR outputs per thread
static_assert(kPerThread * kEnvThreads == kEnvSamples, "tile must split evenly");
constexpr int kXr = 17;                          // exchange row stride (doubles)
constexpr int kXrElems = 16 * kXr;               // 272 doubles per transform

constexpr int kOffQ = 0;                                   // short[4352]    TMA staging
constexpr int kOffC = kOffQ + kEnvSamples * 2;             // double[4352]   continuous FIR output
constexpr int kOffXhead = kOffC + kEnvSamples * 8;         // double[16][16] first 16 inputs of each window
constexpr int kOffHeads = kOffXhead + 16 * 16 * 8;         // double[16][16] zero-history outputs
constexpr int kOffXchg = kOffHeads + 16 * 16 * 8;          // double[16][272] exchange, one component at a time
constexpr int kOffBar = kOffXchg + kEnvH * kXrElems * 8;
constexpr int kEnvSmem = kOffBar + 16;
static_assert(kEnvSmem <= 115712, "two CTAs per SM");

// reference include/bandpass_coeffs.h:1-7 — coeffs[0][0..8]; the filter is symmetric.
__device__ __forceinline__ double fir_tap(int k) {
    constexpr double c[9] = {-0.0023470, 0.0044613, -0.0114627, 0.0226382, -0.0405147,
                             0.0580037,  -0.0779167, 0.0882711, 0.9065095};
    return c[k];
}

// 256-point complex FFT across 16 lanes, exchanging one component at a time through xr
// (16 x 17 doubles). In: v[a] = z[16 a + lane16]. Out: register r holds Z[lane16 + 16 * fft16_out_index(r)].
__device__ __forceinline__ void fft256_split(double2 (&v)[16], int lane16, double *xr, const double2 *tw1, unsigned mask) {
    fft16<double>(v);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int c = fft16_out_index(r);
        if (c != 0) v[r] = cmul<double2>(v[r], tw1[c * 16 + lane16]);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) xr[fft16_out_index(r) * kXr + lane16] = v[r].x;
    __syncwarp(mask);
    double re[16];
#pragma unroll
    for (int b = 0; b < 16; ++b) re[b] = xr[lane16 * kXr + b];
    __syncwarp(mask);
#pragma unroll
    for (int r = 0; r < 16; ++r) xr[fft16_out_index(r) * kXr + lane16] = v[r].y;
    __syncwarp(mask);
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        v[b].x = re[b];
        v[b].y = xr[lane16 * kXr + b];
    }
    __syncwarp(mask);
    fft16<double>(v);
}

// sum_fft of reference src/tempo_atk_sort.c:142-150 for one hop, by the 16 lanes of a half-warp (see
// the header comment). xr[k] = |X_k|^2, k = 0..256. Returns (double)sum_fft on every lane.
__device__ __forceinline__ double float_chain_scan(const double *xr, int lane16, unsigned mask) {
    const int lane_base = (threadIdx.x & 16); // first lane of this half-warp inside its warp
    // bins 0..16 one by one (the sum climbs through several binades here), every lane redundantly
    float sf = 0.0f;
#pragma unroll
    for (int k = 0; k <= 16; ++k) sf = (float)((double)sf + xr[k]);
    double r = (double)sf;
    int kdone = 16; // bins 0..kdone are in r
    int row = 1;    // next row of 16 bins: 16 row + 1 .. 16 row + 16
    while (row < 16) {
        const int rhi = __double2hiint(r), rlo = __double2loint(r);
        const int ex = (rhi >> 20) & 0x7ff;
        if (ex < 1023 - 126 || ex > 1023 + 127) {
            // zero, subnormal, inf or nan: no binade to scan in; one plain reference step
            r = (double)(float)(r + xr[kdone + 1]);
            kdone += 1;
            if (kdone == 16 * row + 16) row += 1;
            continue;
        }
        // binade e = ex - 1023, float grid g = 2^(e-23); r = q g with 2^23 <= q < 2^24
        const int hiM = ((ex + 29) << 20) | 0x80000; // M = 1.5 * 2^(e+29): ulp(M) = g
        const double M = __hiloint2double(hiM, 0);
        int q = ((rhi & 0xFFFFF) << 3) | (int)((unsigned)rlo >> 29) | 0x800000;
        int kstar = 0, qb = 0;
        for (; row < 16; ++row) {
            const int k = 16 * row + 1 + lane16;
            const double t = xr[k] + M; // low word = RN_g(p) / g while p < 2^32 g
            unsigned I = (__double2hiint(t) == hiM) ? (unsigned)__double2loint(t) : (1u << 25);
            I = min(I, 1u << 25);
            if (k <= kdone) I = 0u;
            int incl = (int)I;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const int up = __shfl_up_sync(mask, incl, o, 16);
                if (lane16 >= o) incl += up;
            }
            const int tot = q + incl;
            const unsigned crossed = (__ballot_sync(mask, tot >= (1 << 24)) >> lane_base) & 0xFFFFu;
            if (crossed) {
                const int src = __ffs(crossed) - 1;
                kstar = 16 * row + 1 + src;
                qb = __shfl_sync(mask, tot - (int)I, src, 16); // q + prefix before bin k*
                break;
            }
            q = __shfl_sync(mask, tot, 15, 16);
        }
        const int qq = kstar ? qb : q; // 2^23 <= qq < 2^24: the float sum so far is qq g
        const double s = __hiloint2double((ex << 20) | ((qq & 0x7FFFFF) >> 3), (qq & 7) << 29);
        if (!kstar) return s; // the rest of the chain stayed in this binade
        // bin k* with the reference's own sequence (double add, round to float): next binade
        r = (double)(float)(s + xr[kstar]);
        kdone = kstar;
        if (kstar == 16 * row + 16) row += 1;
    }
    return r;
}
} // namespace

__global__ void __launch_bounds__(kEnvThreads, 2) envelope_kernel(EnvelopeParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const SongDesc sd = p.songs[blockIdx.y];
    const int tile0 = blockIdx.x * kTilesPerCta;
    if (tile0 * kEnvH >= sd.n_hops) return;
    const SongNorm nm = p.norm[blockIdx.y];
    if (nm.status != 0) return;
    const int n_tiles = min(kTilesPerCta, (sd.n_hops - tile0 * kEnvH + kEnvH - 1) / kEnvH);

    short *qs = reinterpret_cast<short *>(smem + kOffQ);
    double *cbuf = reinterpret_cast<double *>(smem + kOffC);
    double *xhead = reinterpret_cast<double *>(smem + kOffXhead);
    double *heads = reinterpret_cast<double *>(smem + kOffHeads);
    double *xchg_all = reinterpret_cast<double *>(smem + kOffXchg);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + kOffBar);

    const int tid = threadIdx.x;
    const short *stream = p.stream + (p.dup ? sd.q_off : sd.pcm_off);
    auto issue_tile = [&](int t) { // one elected thread
        const int m0 = (tile0 + t) * kEnvH;
        const int h_cnt = min(kEnvH, sd.n_hops - m0);
        const long long base = (long long)m0 * kHop;
        const int n_need = (h_cnt + 1) * kHop; // always inside the song: (m + 2) * 256 <= 512 F <= n
        const unsigned bytes = (unsigned)(p.dup ? n_need : 2 * n_need);
        const short *src = stream + (p.dup ? (base >> 1) : base);
        fence_proxy_async();
        mbar_arrive_expect_tx(bar, bytes);
        tma_load_1d(qs, src, bytes, bar);
    };
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        issue_tile(0);
    }
    __syncthreads();

    const int w = tid >> 4, lane16 = tid & 15;
    const unsigned hw_mask = 0xFFFFu << (16 * ((tid >> 4) & 1));
    const double mean_d = nm.mean_d, inv_var_d = nm.inv_var_d;
    const int dup = p.dup;
    unsigned parity = 0;

    for (int t = 0; t < n_tiles; ++t) {
        const int m0 = (tile0 + t) * kEnvH;
        const int h_cnt = min(kEnvH, sd.n_hops - m0);
        mbar_wait(bar, parity);
        parity ^= 1u;

        // ---- normalise + continuous FIR, 17 consecutive outputs per thread
        {
            const int j0 = kPerThread * tid;
            double xv[16 + kPerThread];
#pragma unroll
            for (int i = 0; i < 16 + kPerThread; ++i) {
                const int idx = j0 - 16 + i;
                const int ii = idx < 0 ? 0 : idx;
                // (s / 32768 - mean_d) / var_d, reference src/tempo_atk_sort.c:110-113; the divide is
                // a multiply by the reciprocal (<= 1 ulp apart)
                const double s = int_to_double_exact((int)qs[dup ? (ii >> 1) : ii]);
                const double x = fma(s, 1.0 / 32768, -mean_d) * inv_var_d;
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
        if (tid == 0 && t + 1 < n_tiles) issue_tile(t + 1); // staging buffer is free: prefetch

        // ---- heads: first 16 outputs of windows 1..15 with zero history
        if (w >= 1) {
            const double *xh = xhead + w * 16;
            const int tt = lane16;
            double y = 0;
#pragma unroll
            for (int k = 7; k >= 1; --k) {
                const double a = (tt - k >= 0) ? xh[tt - k] : 0.0;
                const double b = (tt - 16 + k >= 0) ? xh[tt - 16 + k] : 0.0;
                y += fir_tap(k) * (a + b);
            }
            y += ((tt - 8 >= 0) ? xh[tt - 8] : 0.0) * fir_tap(8);
            y += fir_tap(0) * (xh[tt] + 0.0);
            heads[w * 16 + tt] = y;
        }
        __syncthreads();

        // ---- 16 x (512-point double real FFT + float-accumulated power), one hop per half-warp
        if (w < h_cnt) {
            double2 v[16];
#pragma unroll
            for (int a = 0; a < 16; ++a) v[a] = *reinterpret_cast<const double2 *>(cbuf + w * kHop + 32 * a + 2 * lane16);
            if (w >= 1 && lane16 < 8) v[0] = *reinterpret_cast<const double2 *>(heads + w * 16 + 2 * lane16);
            double *xr = xchg_all + w * kXrElems;
            fft256_split(v, lane16, xr, p.tw1, hw_mask);
            // even/odd split needs Z[256 - k]: publish Z one component at a time
            double br[8], bi[8];
#pragma unroll
            for (int r = 0; r < 16; ++r) xr[lane16 + 16 * fft16_out_index(r)] = v[r].x;
            __syncwarp(hw_mask);
#pragma unroll
            for (int d = 0; d < 8; ++d) br[d] = xr[(256 - (lane16 + 16 * d)) & 255];
            __syncwarp(hw_mask);
#pragma unroll
            for (int r = 0; r < 16; ++r) xr[lane16 + 16 * fft16_out_index(r)] = v[r].y;
            __syncwarp(hw_mask);
#pragma unroll
            for (int d = 0; d < 8; ++d) bi[d] = xr[(256 - (lane16 + 16 * d)) & 255];
            __syncwarp(hw_mask);
            // |X_k|^2, k = 0..256, into the same buffer
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const int k = lane16 + 16 * d;
                const double2 A = v[fft16_reg_of(d)];
                if (k == 0) {
                    const double x0 = A.x + A.y, xn = A.x - A.y; // X_0 and X_256 are real
                    xr[0] = x0 * x0;
                    xr[256] = xn * xn;
                } else {
                    const double2 wk = p.tw2[k];
                    const double sr = A.x + br[d], si = A.y - bi[d];
                    const double dr = A.x - br[d], di = A.y + bi[d];
                    const double tr = dr * wk.x - di * wk.y;
                    const double ti = dr * wk.y + di * wk.x;
                    const double ar = sr + ti, ai = si - tr;
                    const double cr = sr - ti, ci = si + tr;
                    xr[k] = 0.25 * (ar * ar + ai * ai);
                    xr[256 - k] = 0.25 * (cr * cr + ci * ci);
                }
            }
            if (lane16 == 0) {
                const double2 A = v[fft16_reg_of(8)];
                xr[128] = A.x * A.x + A.y * A.y;
            }
            __syncwarp(hw_mask);
            const double e = float_chain_scan(xr, lane16, hw_mask);
            if (lane16 == 0) p.energy[sd.env_off + m0 + w] = e;
        }
        __syncthreads(); // cbuf / heads / xchg free for the next tile
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
    const int per_cta = kEnvH * kTilesPerCta;
    dim3 grid((unsigned)((max_hops + per_cta - 1) / per_cta), (unsigned)n_songs);
    envelope_kernel<<<grid, kEnvThreads, kEnvSmem, st>>>(p);
    return cudaGetLastError();
}

} // namespace blx
