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
// (fft16.cuh, double), and the even/odd split leaves |X_k|^2 of every hop in shared memory.
//
// The float accumulation is a chain of 257 dependent roundings per hop: latency-, not
// throughput-bound, and only 16 chains exist per tile. The CTA is therefore warp-specialised:
//   warps 0..7  (producers)  TMA tile -> FIR -> heads -> 16 FFTs -> P[16][257]
//   warp  8     (consumer)   lane h replays the chain of hop h of the PREVIOUS tile
// so the chains of tile t overlap the FIR/FFT work of tile t+1 (named barriers 2/3 hand the P buffer
// back and forth; the int16 tile of t+1 is fetched by one 1-D bulk copy, cp.async.bulk / UBLKCP, as
// soon as the FIR of tile t has consumed the staging buffer). A CTA walks kTilesPerCta tiles.
//
// The chain itself: s <- (float)((double)s + p). While s stays inside one binade [2^e, 2^(e+1)) the
// float grid is 2^(e-23), which is exactly the double grid of the binade [2^(e+29), 2^(e+30)). With
// C = 2^(e+29) the value S' = s + C is exact and S' + p is ONE double addition whose IEEE rounding
// (nearest, ties to even, same parity) is the rounding to the float grid; s + p >= 2^(e+1) shows as
// bits(S' + p) >= bits(C + 2^(e+1)), and that step is redone through the reference's own
// double-add / float-convert sequence and re-based. The result equals the reference's except when
// s + p lies within 2^-53 (relative) of a float rounding midpoint (the reference rounds twice).
//
// FP64 throughout; the summation order of the FIR is the reference's. FMA contraction is allowed
// here (the reference has none): it perturbs E[m] by ~1e-16 relative, nine orders of magnitude
// below the 1e-7 onset-count cliff measured in SURVEY.md App. B.
#include "blx_common.cuh"
#include "fft16.cuh"
#include "kernels.h"

namespace blx {

namespace {
constexpr int kEnvProducers = 256;               // 8 warps
constexpr int kEnvThreads = kEnvProducers + 32;  // + 1 consumer warp
constexpr int kEnvH = 16;                        // hops per tile
constexpr int kTilesPerCta = 8;
constexpr int kEnvSamples = (kEnvH + 1) * kHop;  // 4352 stream samples per tile
constexpr int kPerThread = kEnvSamples / kEnvProducers; // 17 consecutive FIR outputs per thread
static_assert(kPerThread * kEnvProducers == kEnvSamples, "tile must split evenly");
constexpr int kXr = 17;                          // exchange row stride (doubles)
constexpr int kXrElems = 16 * kXr;               // 272 doubles per transform (>= 257 for the split)

constexpr int kOffQ = 0;                                   // short[4352]    TMA staging
constexpr int kOffC = kOffQ + kEnvSamples * 2;             // double[4352]   continuous FIR output
constexpr int kOffXhead = kOffC + kEnvSamples * 8;         // double[16][16] first 16 inputs of each window
constexpr int kOffHeads = kOffXhead + 16 * 16 * 8;         // double[16][16] zero-history outputs
constexpr int kOffXchg = kOffHeads + 16 * 16 * 8;          // double[16][272] one component at a time
constexpr int kOffP = kOffXchg + kEnvH * kXrElems * 8;     // double[16][257] |X_k|^2 of the tile
constexpr int kOffBar = kOffP + kEnvH * 257 * 8;
constexpr int kEnvSmem = kOffBar + 16;
static_assert(kEnvSmem <= 115712, "two CTAs per SM");

// reference include/bandpass_coeffs.h:1-7 — coeffs[0][0..8]; the filter is symmetric.
__device__ __forceinline__ double fir_tap(int k) {
    constexpr double c[9] = {-0.0023470, 0.0044613, -0.0114627, 0.0226382, -0.0405147,
                             0.0580037,  -0.0779167, 0.0882711, 0.9065095};
    return c[k];
}

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

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

// State of one accumulation chain (see the header comment).
struct Chain {
    double Sp;   // s + C
    double C;    // 2^(e+29), or 0 when s is zero / outside the normal float range (every step is redone)
    long long L; // bits(C + 2^(e+1)): first value that leaves the binade
    __device__ __forceinline__ void rebase(double r) { // r >= 0 holds a float value
        const int ex = (__double2hiint(r) >> 20) & 0x7ff;
        if (ex >= 1023 - 126 && ex <= 1023 + 127) {
            const int chi = (ex + 29) << 20;
            C = __hiloint2double(chi, 0);
            Sp = r + C; // exact
            L = ((long long)chi << 32) + (1ll << 24);
        } else {
            C = 0.0;
            Sp = r;
            L = 0;
        }
    }
    __device__ __forceinline__ void add(double p) {
        const double A = Sp + p;
        if (__double_as_longlong(A) < L) {
            Sp = A;
        } else { // reference src/tempo_atk_sort.c:147 verbatim, then a new binade
            const double s = Sp - C;
            rebase((double)(float)(s + p));
        }
    }
    __device__ __forceinline__ double value() const { return Sp - C; }
};
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
    double *Pall = reinterpret_cast<double *>(smem + kOffP);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + kOffBar);

    const int tid = threadIdx.x;

    // ============================================================ consumer warp
    if (tid >= kEnvProducers) {
        const int lane = tid - kEnvProducers;
        for (int t = 0; t < n_tiles; ++t) {
            const int m0 = (tile0 + t) * kEnvH;
            const int h_cnt = min(kEnvH, sd.n_hops - m0);
            bar_sync(2, kEnvThreads); // P of tile t is complete
            if (lane < h_cnt) {
                const double *Pw = Pall + lane * 257;
                Chain ch;
                ch.rebase((double)(float)Pw[0]); // sum_fft = (float)(0 + P[0])
#pragma unroll 4
                for (int k = 1; k <= 256; ++k) ch.add(Pw[k]);
                p.energy[sd.env_off + m0 + lane] = ch.value();
            }
            if (t + 1 < n_tiles) bar_arrive(3, kEnvThreads); // P may be overwritten
        }
        return;
    }

    // ============================================================ producer warps
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
    bar_sync(1, kEnvProducers);

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
        bar_sync(1, kEnvProducers);
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
        bar_sync(1, kEnvProducers);

        // ---- 16 x 512-point double real FFT, one per half-warp
        const bool active = w < h_cnt;
        double pk[8], pq[8], p128 = 0.0; // |X_k|^2 for k = lane16 + 16 d, and for 256 - k
        if (active) {
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
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const int k = lane16 + 16 * d;
                const double2 A = v[fft16_reg_of(d)];
                if (k == 0) {
                    const double x0 = A.x + A.y, xn = A.x - A.y; // X_0 and X_256 are real
                    pk[d] = x0 * x0;
                    pq[d] = xn * xn;
                } else {
                    const double2 wk = p.tw2[k];
                    const double sr = A.x + br[d], si = A.y - bi[d];
                    const double dr = A.x - br[d], di = A.y + bi[d];
                    const double tr = dr * wk.x - di * wk.y;
                    const double ti = dr * wk.y + di * wk.x;
                    const double ar = sr + ti, ai = si - tr;
                    const double cr = sr - ti, ci = si + tr;
                    pk[d] = 0.25 * (ar * ar + ai * ai);
                    pq[d] = 0.25 * (cr * cr + ci * ci);
                }
            }
            if (lane16 == 0) {
                const double2 A = v[fft16_reg_of(8)];
                p128 = A.x * A.x + A.y * A.y;
            }
        }
        if (t > 0) bar_sync(3, kEnvThreads); // the consumer has finished with the previous tile's P
        if (active) {
            double *P = Pall + w * 257;
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const int k = lane16 + 16 * d;
                P[k] = pk[d];
                P[256 - k] = pq[d];
            }
            if (lane16 == 0) P[128] = p128;
        }
        __threadfence_block();
        bar_arrive(2, kEnvThreads); // hand P to the consumer
        // all producers are past their cbuf / heads reads before the next FIR overwrites them
        bar_sync(1, kEnvProducers);
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
