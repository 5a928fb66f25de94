// epilogue.cu — the per-song scalar stages, in the reference's operation order.
// Compiled with -fmad=false: the reference is built -std=c99 (no FMA contraction, reference
// CMakeLists.txt:22), and everything here is latency- not throughput-bound.
//
//   epilogue_kernel (BLX_K_EPILOGUE), one CTA per song:
//     - sums the song's partial spectra in a fixed order, then steps 6-9 of the frequency rating
//       (reference src/frequency_sort.c:97-139);
//     - 301 passes of the 7-tap [1,3,6,7,6,3,1]/27 smoothing on the 3807 histogram bins that can
//       reach the integration window, with the reference's float/double rounding sequence, then
//       normalisation and the window integral (reference src/amplitude_sort.c:41-79). Bit-exact
//       with the reference for the same histogram;
//     - bl_mean / bl_variance from the integer sums (reference src/helpers.c:30-49) and the
//       normalisation constants of reference src/tempo_atk_sort.c:105-107.
//   tail_kernel (BLX_K_TAIL), one thread per song: reference src/tempo_atk_sort.c:184-287 as a single
//     streaming pass (log compression, x2 zero-stuffing, 6th-order IIR, rectified difference, weighted
//     mix, two width-19 running-sum filters with their edge quirks, onset count) + the rating of
//     reference src/analyze.c:63-79.
//   rect_filter_kernel: bl_rectangular_filter (reference src/tempo_atk_sort.c:19-40) for the public
//     helper of bliss.h.
#include <math.h>

#include "blx_common.cuh"
#include "kernels.h"

namespace blx {

namespace {
constexpr int kEpThreads = 256;
constexpr int kSmW = kHistBins + 6; // 3 zero bins of padding on each side
constexpr int kEpRun = 15;          // consecutive bins per thread in the smoothing passes (256 x 15 >= 3807)
static_assert(kEpRun * kEpThreads >= kHistBins, "every bin has an owner");

__device__ __forceinline__ float block_max(float v, float *scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = scratch[0];
    for (int i = 1; i < kEpThreads / 32; ++i) r = fmaxf(r, scratch[i]);
    __syncthreads();
    return r;
}
} // namespace

__global__ void __launch_bounds__(kEpThreads) epilogue_kernel(EpilogueParams p) {
    __shared__ float ps[257];
    __shared__ float hA[kSmW + kEpRun], hB[kSmW + kEpRun]; // + read-ahead of the last owner's window
    __shared__ float scratch[8];
    __shared__ int sh_status;
    const int s = blockIdx.x;
    const int tid = threadIdx.x;
    const SongDesc sd = p.songs[s];
    SongNorm out;
    out.mean_d = 0.0; out.inv_var_d = 0.0; out.amplitude = 0.0f; out.frequency = 0.0f; out.status = 0;
    out.mean = 0; out.variance = 0; out.pad = 0;

    // ------------------------------------------------------------ frequency (A.1 steps 6-9)
    if (p.what & BLX_DO_FREQUENCY) {
        float acc = 0.0f;
        if (tid >= 1) // bins 1..255; ps[256] stays 0 (Nyquist ignored, reference src/frequency_sort.c:58-62,88)
            for (int q = 0; q < sd.n_parts; ++q) acc += p.partials[(size_t)(sd.part_off + q) * 256 + tid];
        // d = tid (1..255) and d = 256 handled by thread 0
        const int d = (tid == 0) ? 256 : tid;
        float v = (float)sqrt((double)(acc / 512.0f));
        const float peak = block_max(v, scratch);
        v = (float)(20 * log10((double)(v / peak)) - 3);
        ps[d] = v;
        __syncthreads();
        if (tid == 0) {
            float b0 = (ps[2] + ps[4]) / 2;
            float b1 = (ps[6] + ps[8]) / 2;
            float b2 = 0.0f, b3 = 0.0f, b4 = 0.0f;
            for (int i = 10; i <= 60; ++i) b2 += ps[i];
            b2 /= 50;
            for (int i = 61; i <= 118; ++i) b3 += ps[i];
            b3 /= 57;
            for (int i = 119; i <= 234; ++i) b4 += ps[i];
            b4 /= 115;
            const float bands_sum = b4 + b3 + b2 - b0 - b1;
            out.frequency = (float)((1. / 3.) * (double)bands_sum + 68. / 3.);
        }
    }

    // ------------------------------------------------------------ amplitude (A.2 steps 3-5)
    const SongStats st = p.stats ? p.stats[s] : SongStats{0, 0ull, 0u, 0u};
    const int first_nz = (int)(0x7fffffffu - st.first_inv), last_nz = (int)st.last_p1 - 1;
    if (p.what & BLX_DO_AMPLITUDE) {
        if (last_nz < 0) {
            out.status |= BLX_SONG_SILENT;
            out.amplitude = nanf("");
        } else {
            // Pass 1 counted every sample; the reference skips the zeros in front of the first and behind the
            // last non-zero sample (reference src/amplitude_sort.c:26-39), and its float counters stop at 2^24
            // (x + 1 == x from there on).
            const unsigned *gh = p.hist + (size_t)s * kHistStride;
            const unsigned trimmed = (unsigned)first_nz + (unsigned)(sd.n_samples - 1 - last_nz);
            for (int i = tid; i < kSmW + kEpRun; i += kEpThreads) {
                const int b = i - 3;
                unsigned c = (b >= 0 && b < kHistBins) ? gh[b] : 0u;
                if (b == 32768 - kHistLo) c -= trimmed;
                hA[i] = (float)min(c, 1u << 24);
                hB[i] = 0.0f;
            }
            __syncthreads();
            float *h = hA, *sm = hB;
            // every thread owns kEpRun consecutive bins: 21 loads feed its 15 outputs (a sliding window in
            // registers); a stride of 15 words between threads is bank-conflict free
            const int i0 = 3 + kEpRun * tid;
            for (int g = 0; g < kSmoothPasses; ++g) {
                if (i0 < 3 + kHistBins) {
                    float w[kEpRun + 6];
#pragma unroll
                    for (int k = 0; k < kEpRun + 6; ++k) w[k] = h[i0 - 3 + k];
#pragma unroll
                    for (int k = 0; k < kEpRun; ++k) {
                        const float taps = w[k] + (3 * w[k + 1]) + (6 * w[k + 2]) + (7 * w[k + 3]) + (6 * w[k + 4]) +
                                           (3 * w[k + 5]) + w[k + 6];
                        if (i0 + k < 3 + kHistBins) sm[i0 + k] = (float)(1. / 27. * (double)taps);
                    }
                }
                __syncthreads();
                float *t = h; h = sm; sm = t; // h now holds this pass's output (reference copies it back)
            }
            // h = histogram_smooth after the last pass; normalise + integrate in index order
            if (tid == 0) {
                const float span = (float)(first_nz - last_nz);
                float integral = 0.0f;
                for (int b = kIntLo; b <= kIntHi; ++b) {
                    float v = h[b - kHistLo + 3] / span;
                    v = (float)((double)v * 100.);
                    v = fabsf(v);
                    integral += v;
                }
                out.amplitude = -0.2f * integral + 6.0f;
            }
        }
    }

    // ------------------------------------------------------------ statistics (A.3 step 1-2)
    if (tid == 0) {
        if (p.what & BLX_DO_ENVELOPE) {
            const int n = sd.n_samples;
            if (n < 3 * kWin || sd.duration == 0 || 4 * sd.F < 40) out.status |= BLX_SONG_TOO_SHORT;
            if (last_nz < 0) out.status |= BLX_SONG_SILENT;
            if (n > 0) {
                // `int` accumulator of the reference wraps modulo 2^32 in practice (UB in C)
                const int mean = (int)(unsigned)(unsigned long long)st.sum / n;
                // sum (s - m)^2 = sum s^2 - 2 m sum s + n m^2, exact in 64-bit integers
                const long long dev = (long long)st.sumsq - 2ll * mean * st.sum + (long long)n * mean * mean;
                const int var = (int)(dev / n);
                if (var == 0) out.status |= BLX_SONG_FLAT;
                out.mean = mean;
                out.variance = var;
                out.mean_d = (double)mean / 32768;
                double var_d = (double)var / 32768;
                var_d /= 32768;
                out.inv_var_d = 1.0 / var_d;
            }
        }
        p.norm[s] = out;
        if (p.frequency) p.frequency[s] = out.frequency;
        sh_status = out.status;
    }
    // E[M], E[M + 1] stay 0 in the reference (calloc, reference src/tempo_atk_sort.c:84): the envelope kernel
    // writes hops 0..M-1 only. A song it skips (status != 0) gets a cleared row.
    if (p.energy && (p.what & BLX_DO_ENVELOPE)) {
        __syncthreads();
        double *row = p.energy + sd.env_off;
        const int nb = (max(2 * sd.F, 2) + 7) & ~7; // the whole row incl. its padding (plan_songs): logcomp reads all of it
        for (int i = (sh_status ? 0 : max(sd.n_hops, 0)) + tid; i < nb; i += kEpThreads) row[i] = 0.0;
    }
}

cudaError_t launch_epilogue(const EpilogueParams &p, int n_songs, cudaStream_t st) {
    epilogue_kernel<<<n_songs, kEpThreads, 0, st>>>(p);
    return cudaGetLastError();
}

// =====================================================================================
// tail
// =====================================================================================
namespace {
constexpr int kTailSongs = 32;   // songs per CTA, one per lane
constexpr int kTailThreads = 64; // warp 0: a song's owner (everything but the IIR in steady state), warp 1: its IIR
constexpr int kBox = 19;
// The IIR recurrence is a 7-operation dependent chain per sample (~92 cycles) and the rest of a sample's work
// (~45 FP64 operations) does not feed back into it, so in steady state the two run on two warps (two
// schedulers, two FP64 pipes): warp 1 streams y[n] through a shared-memory ring of kTailBufs blocks of
// kTailBlk sample pairs, warp 0 consumes them. Named barriers 1.. (full) and 1 + kTailBufs.. (empty).
constexpr int kTailBlk = 8;      // input pairs (2 samples each) per ring block
constexpr int kTailBufs = 4;
// the two hand-overs between the owner and the IIR warp are named barriers as well, one warp arriving and the other
// waiting: the warps reach them from different places in the code, which __syncthreads() (one call site for the whole
// block) does not allow
constexpr int kBarPublished = 1 + 2 * kTailBufs, kBarHandedBack = 2 + 2 * kTailBufs;

// bar.sync / bar.arrive are warp-ALIGNED instructions: the whole warp has to execute them together. Both are used behind
// per-lane `if (run_env)` blocks (a lane without a song idles), so the warp is re-converged first (compute-sanitizer
// --tool synccheck flags the barrier otherwise; tools/sanitize.sh).
__device__ __forceinline__ void bar_sync(int id) {
    __syncwarp();
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(kTailThreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id) {
    __syncwarp();
    __threadfence_block(); // the ring block / its release is visible before the other warp passes its bar.sync
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(kTailThreads) : "memory");
}

// reference include/bandpass_coeffs.h:484-492 (data literals)
__constant__ double c_lp_b[7] = {1.9510e-05, 1.1706e-04, 2.9266e-04, 3.9021e-04, 2.9266e-04, 1.1706e-04, 1.9510e-05};
__constant__ double c_lp_a[7] = {1.00000, -4.59007, 8.91034, -9.34191, 5.56998, -1.78845, 0.24136};

// x / d for a compile-time constant d, correctly rounded: q = RN(x * (1/d)), one FMA residual, one FMA
// correction (Markstein). Replaces the ~25-instruction IEEE division in the per-sample loop; the
// rounding is the division's own for every finite x that is not within a few ulps of the subnormal or
// overflow range (the envelope samples are O(1)).
template <int D> __device__ __forceinline__ double div_const(double x) {
    const double r = 1.0 / (double)D; // RN(1/D), folded at compile time
    const double q = x * r;
    const double rem = fma(-(double)D, q, x);
    return fma(rem, r, q);
}

struct PeakCounter { // onsets of reference src/tempo_atk_sort.c:277-280, fed one sample at a time
    double prev2, prev1;
    int index_next; // index of the next sample to be pushed
    int beat;
    int n2;
    double eps;
    __device__ void push(double s) {
        // centre = index_next - 1, valid for 1 <= centre <= n2 - 2
        const int centre = index_next - 1;
        if (centre >= 1 && centre <= n2 - 2)
            if (((prev1 - prev2) > eps) && ((prev1 - s) > eps)) beat++;
        prev2 = prev1;
        prev1 = s;
        index_next++;
    }
};
} // namespace

__global__ void __launch_bounds__(kTailThreads) tail_kernel(TailParams p, int n_songs) {
    __shared__ double ring1[kBox][kTailSongs]; // last 19 inputs of the first box filter (ss)
    __shared__ double ring2[kBox][kTailSongs]; // last 19 inputs of the second box filter
    __shared__ double ybuf[kTailBufs][2 * kTailBlk][kTailSongs]; // y[n] from the IIR warp to the owner warp
    __shared__ double hand[9][kTailSongs];     // y1..y6, e1..e3 handed to the IIR warp and back
    __shared__ int sh_blocks;                  // ring blocks of the piped phase: the same for all songs of the CTA
    const int tx = threadIdx.x & 31;
    const int role = threadIdx.x >> 5; // 0 owner, 1 IIR
    const int s = blockIdx.x * kTailSongs + tx;
    const bool in_range = s < n_songs;
    const SongDesc sd = p.songs[in_range ? s : 0];
    const SongNorm nm = p.norm[in_range ? s : 0];
    const bool run_env = in_range && (p.what & BLX_DO_ENVELOPE) && nm.status == 0;

    if (role == 1) {
        // ---- IIR warp: y[n] for the piped phase (reference src/tempo_atk_sort.c:201-218, same operation order)
        bar_sync(kBarPublished); // the owners have run their start-up samples and published state + block count
        const int n_blocks = sh_blocks;
        const double *X = p.xlog + sd.env_off;
        double y1 = 0, y2 = 0, y3 = 0, y4 = 0, y5 = 0, y6 = 0, e1 = 0, e2 = 0, e3 = 0;
        if (run_env) {
            y1 = hand[0][tx]; y2 = hand[1][tx]; y3 = hand[2][tx]; y4 = hand[3][tx]; y5 = hand[4][tx]; y6 = hand[5][tx];
            e1 = hand[6][tx]; e2 = hand[7][tx]; e3 = hand[8][tx];
        }
        int i = 16; // input index of sample j = 32
        double cur[kTailBlk];
#pragma unroll
        for (int k = 0; k < kTailBlk; ++k) cur[k] = (run_env && n_blocks > 0) ? X[i + k] : 0.0;
        for (int b = 0; b < n_blocks; ++b) {
            const int slot = b % kTailBufs;
            bar_sync(1 + kTailBufs + slot); // the owner has released this ring block
            if (run_env) {
                double nxt[kTailBlk]; // next block's inputs: a block is ~1500 cycles of chain, the loads land meanwhile
#pragma unroll
                for (int k = 0; k < kTailBlk; ++k) nxt[k] = X[i + kTailBlk + k]; // rows carry 16 doubles of slack
                asm volatile("prefetch.global.L1 [%0];" ::"l"(X + i + 4 * kTailBlk));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(X + i + 4 * kTailBlk + 4));
#pragma unroll
                for (int k = 0; k < kTailBlk; ++k) {
                    const double e0 = cur[k];
#pragma unroll
                    for (int odd = 0; odd < 2; ++odd) {
                        double d;
                        if (!odd) { d = c_lp_b[0] * e0; d += c_lp_b[2] * e1; d += c_lp_b[4] * e2; d += c_lp_b[6] * e3; }
                        else { d = c_lp_b[1] * e0; d += c_lp_b[3] * e1; d += c_lp_b[5] * e2; }
                        double c = c_lp_a[1] * y1;
                        c += c_lp_a[2] * y2; c += c_lp_a[3] * y3; c += c_lp_a[4] * y4; c += c_lp_a[5] * y5; c += c_lp_a[6] * y6;
                        const double y = d - c;
                        y6 = y5; y5 = y4; y4 = y3; y3 = y2; y2 = y1; y1 = y;
                        ybuf[slot][2 * k + odd][tx] = y;
                    }
                    e3 = e2; e2 = e1; e1 = e0;
                }
#pragma unroll
                for (int k = 0; k < kTailBlk; ++k) cur[k] = nxt[k];
                i += kTailBlk;
            }
            bar_arrive(1 + slot); // block full
        }
        if (run_env) {
            hand[0][tx] = y1; hand[1][tx] = y2; hand[2][tx] = y3; hand[3][tx] = y4; hand[4][tx] = y5; hand[5][tx] = y6;
            hand[6][tx] = e1; hand[7][tx] = e2; hand[8][tx] = e3;
        }
        bar_arrive(kBarHandedBack); // state handed back (one-way: the owner warp waits for it)
        return;
    }

    blx_result res;
    res.tempo = 0.0f; res.attack = 0.0f; res.amplitude = nm.amplitude; res.frequency = nm.frequency;
    res.force = 0.0f; res.calm_or_loud = 2; res.beat = 0; res.status = nm.status;

    // State of the owner's song lives at function scope: the barrier operations below must run in warp-uniform
    // control flow, with the per-lane work (lanes without envelope work idle) nested inside.
    {
        const double *X = p.xlog + sd.env_off; // log(1 + mu E) / log(1 + mu), from logcomp_kernel
        const int nb = 2 * sd.F;
        const int n2 = 2 * nb;
        const float lambda = 0.8f;
        const double w_lp = (double)(1 - lambda);
        const double w_df = (double)(lambda * 172);
        const double eps = (double)0.000001f;

        double x1 = 0, x2 = 0, x3 = 0, x4 = 0, x5 = 0, x6 = 0; // t1 history
        double y1 = 0, y2 = 0, y3 = 0, y4 = 0, y5 = 0, y6 = 0; // t2 history
        double atk_sum = 0;
        double ts1 = 0, ts2 = 0;
        double wa_last = 0;
        int r1 = 0, r2 = 0; // ring write positions (= index mod 19)
        PeakCounter pk;
        pk.prev2 = 0; pk.prev1 = 0; pk.index_next = 9; pk.beat = 0; pk.n2 = n2; pk.eps = eps;
        // out2[0..8] = 0: the counter starts as if samples 0..8 (all zero) had been pushed.

        // consume one input of the second box filter, in index order p2 = 0, 1, 2, ...
        int p2 = 0;
        auto feed2 = [&](double v) {
            if (p2 < kBox) {
                ts2 += v;
            } else {
                pk.push(div_const<kBox>(ts2)); // out2[p2 - 10]
                ts2 -= ring2[r2][tx]; // in2[p2 - 19]
                ts2 += v;
            }
            ring2[r2][tx] = v;
            r2 = (r2 + 1 == kBox) ? 0 : r2 + 1;
            p2++;
        };

        // One sample of the chain with every start-up / last-sample condition spelled out.
        auto step_generic = [&](int j, double x0) {
            // step 7: IIR (reference src/tempo_atk_sort.c:201-218)
            double d = 0, c = 0;
            d += c_lp_b[0] * x0; d += c_lp_b[1] * x1; d += c_lp_b[2] * x2; d += c_lp_b[3] * x3;
            d += c_lp_b[4] * x4; d += c_lp_b[5] * x5; d += c_lp_b[6] * x6;
            c += c_lp_a[1] * y1; c += c_lp_a[2] * y2; c += c_lp_a[3] * y3;
            c += c_lp_a[4] * y4; c += c_lp_a[5] * y5; c += c_lp_a[6] * y6;
            const double y = d - c; // / buttera[0], which is 1.0 (reference include/bandpass_coeffs.h:489)
            // step 8: rectified difference (reference src/tempo_atk_sort.c:221-226)
            double df;
            if (j == 0) df = y;
            else { df = y - y1; df = (df > 0) ? df : 0; }
            // step 9: weighted mix (reference src/tempo_atk_sort.c:229-232)
            const double wa = w_lp * y + div_const<10>(w_df * df);
            x6 = x5; x5 = x4; x4 = x3; x3 = x2; x2 = x1; x1 = x0;
            y6 = y5; y5 = y4; y4 = y3; y3 = y2; y2 = y1; y1 = y;
            // step 10: attack sum and the ss array (reference src/tempo_atk_sort.c:246-263)
            double in1;
            if (j < n2 - 1) { atk_sum += wa; in1 = wa; }
            else { in1 = 0; wa_last = wa; }
            // step 11a: first running-sum filter (out = wa array, in = ss)
            if (j < kBox) ts1 += in1;
            if (j >= 10) {
                const int q = j - 10; // index of the filter-1 output becoming final now
                double o1;
                if (q <= 8) o1 = div_const<kBox>(ring1[q][tx]); // untouched entries keep wa[q], then /19
                else o1 = div_const<kBox>(ts1);                 // value before this step's update
                feed2(o1);
            }
            if (j >= kBox) {
                ts1 -= ring1[r1][tx]; // in1[j - 19]
                ts1 += in1;
            }
            ring1[r1][tx] = in1;
            r1 = (r1 + 1 == kBox) ? 0 : r1 + 1;
        };
        // The same sample for 29 <= j < n2 - 1, where both filters are warm and no edge applies:
        // branch-free, so everything but the IIR recurrence overlaps it. The x history is zero-stuffed
        // (x1 = x3 = x5 = 0 on even samples, x0 = x2 = x4 = x6 = 0 on odd ones); the skipped products
        // are exact zeros, so the partial sums are the reference's.
        // everything of a steady-state sample but the IIR: y is this sample's filter output, y1 the previous one
        auto step_rest = [&](double y) {
            double df = y - y1;
            df = (df > 0) ? df : 0;
            const double wa = w_lp * y + div_const<10>(w_df * df);
            y1 = y;
            atk_sum += wa;
            const double o1 = div_const<kBox>(ts1);
            const double s2 = div_const<kBox>(ts2); // out2[p2 - 10]
            if (((pk.prev1 - pk.prev2) > eps) && ((pk.prev1 - s2) > eps)) pk.beat++;
            pk.prev2 = pk.prev1; pk.prev1 = s2; pk.index_next++;
            ts2 -= ring2[r2][tx]; ts2 += o1;
            ring2[r2][tx] = o1;
            r2 = (r2 + 1 == kBox) ? 0 : r2 + 1;
            p2++;
            ts1 -= ring1[r1][tx]; ts1 += wa;
            ring1[r1][tx] = wa;
            r1 = (r1 + 1 == kBox) ? 0 : r1 + 1;
        };
        auto step_steady = [&](double d) {
            double c = c_lp_a[1] * y1;
            c += c_lp_a[2] * y2; c += c_lp_a[3] * y3; c += c_lp_a[4] * y4; c += c_lp_a[5] * y5; c += c_lp_a[6] * y6;
            const double y = d - c;
            const double yprev = y1;
            step_rest(y); // sets y1 = y
            y6 = y5; y5 = y4; y4 = y3; y3 = y2; y2 = yprev;
        };

        int j = 0;
        if (run_env)
            for (; j < 32; ++j) step_generic(j, (j & 1) ? 0.0 : X[j >> 1]); // n2 >= 40 (BLX_SONG_TOO_SHORT otherwise)
        // e1, e2, e3: the three most recent even-index inputs (x at j-2, j-4, j-6 for even j)
        double e1 = x2, e2 = x4, e3 = x6;
        // ---- piped phase: whole ring blocks that EVERY song of the CTA still has in its branch-free range
        // (pairs j, j + 1 with j + 2 <= n2 - 1), the IIR on warp 1
        {
            const int my_blocks = run_env ? ((n2 - 1 - 32) / 2) / kTailBlk : 0x7fffffff;
            const int nblk = __reduce_min_sync(0xffffffffu, my_blocks);
            if (tx == 0) sh_blocks = (nblk == 0x7fffffff) ? 0 : nblk;
            if (run_env) {
                hand[0][tx] = y1; hand[1][tx] = y2; hand[2][tx] = y3; hand[3][tx] = y4; hand[4][tx] = y5; hand[5][tx] = y6;
                hand[6][tx] = e1; hand[7][tx] = e2; hand[8][tx] = e3;
            }
            bar_arrive(kBarPublished); // one-way: the IIR warp waits for it (bar.sync)
            const int n_blocks = (nblk == 0x7fffffff) ? 0 : nblk; // what lane 0 just published in sh_blocks
            for (int k = 0; k < kTailBufs && k < n_blocks; ++k) bar_arrive(1 + kTailBufs + k); // the ring starts empty
            for (int b = 0; b < n_blocks; ++b) {
                const int slot = b % kTailBufs;
                bar_sync(1 + slot); // block full
                if (run_env) {
#pragma unroll 4
                    for (int k = 0; k < 2 * kTailBlk; ++k) step_rest(ybuf[slot][k][tx]);
                }
                if (b + kTailBufs < n_blocks) bar_arrive(1 + kTailBufs + slot); // released (only if it is needed again)
            }
            bar_sync(kBarHandedBack); // the IIR warp has handed its state back
            if (run_env && n_blocks > 0) {
                y1 = hand[0][tx]; y2 = hand[1][tx]; y3 = hand[2][tx]; y4 = hand[3][tx]; y5 = hand[4][tx]; y6 = hand[5][tx];
                e1 = hand[6][tx]; e2 = hand[7][tx]; e3 = hand[8][tx];
                j += 2 * kTailBlk * n_blocks;
            }
        }
        if (run_env) {
            int i = j >> 1;
            // one even + one odd sample from the even-index input e0
            auto pair = [&](double e0) {
                double d = c_lp_b[0] * e0; // even sample
                d += c_lp_b[2] * e1; d += c_lp_b[4] * e2; d += c_lp_b[6] * e3;
                step_steady(d);
                d = c_lp_b[1] * e0; // odd sample
                d += c_lp_b[3] * e1; d += c_lp_b[5] * e2;
                step_steady(d);
                e3 = e2; e2 = e1; e1 = e0;
            };
            // read-ahead of four to eight inputs (8-16 samples, >= 1000 cycles) in two register sets that
            // take turns, so that no load has to be waited for (or its value moved) before its turn
            // (rows carry 16 doubles of slack)
            double q0 = X[i], q1 = X[i + 1], q2 = X[i + 2], q3 = X[i + 3];
            for (; j + 16 <= n2 - 1; j += 16, i += 8) {
                // every lane streams its own row (one 32-byte sector per 4 inputs, a DRAM page of its own):
                // pull the sectors of ~50 samples ahead into L1 so that the loads below are L1 hits
                if (2 * (i + 32) < n2) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(X + i + 28));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(X + i + 32));
                }
                const double u0 = X[i + 4], u1 = X[i + 5], u2 = X[i + 6], u3 = X[i + 7];
                pair(q0); pair(q1); pair(q2); pair(q3);
                q0 = X[i + 8]; q1 = X[i + 9]; q2 = X[i + 10]; q3 = X[i + 11];
                pair(u0); pair(u1); pair(u2); pair(u3);
            }
            for (; j + 8 <= n2 - 1; j += 8, i += 4) {
                double e0 = q0; q0 = X[i + 4]; pair(e0);
                e0 = q1; q1 = X[i + 5]; pair(e0);
                e0 = q2; q2 = X[i + 6]; pair(e0);
                e0 = q3; q3 = X[i + 7]; pair(e0);
            }
            for (; j + 2 <= n2 - 1; j += 2, ++i) pair(X[i]);
            // back to the generic history: j is even, x1 = 0, x2 = e1, ...
            x1 = 0; x2 = e1; x3 = 0; x4 = e2; x5 = 0; x6 = e3;
        }
        if (run_env) {
        for (; j < n2; ++j) step_generic(j, (j & 1) ? 0.0 : X[j >> 1]);
        // ---- end quirks of filter 1 (reference src/tempo_atk_sort.c:34-39): indices n2-10 .. n2-1
        {
            // ring1 holds in1[n2-19 .. n2-1]; the oldest is at r1
            double o = ring1[(r1 + 9) % kBox][tx]; // wa[n2 - 10] (== in1 there)
            for (int k = 0; k < kBox; ++k) o += ring1[(r1 + k) % kBox][tx];
            feed2(div_const<kBox>(o));
            for (int q = n2 - 9; q < n2; ++q) {
                const double w = (q == n2 - 1) ? wa_last : ring1[(r1 + (q - (n2 - kBox))) % kBox][tx];
                feed2(div_const<kBox>(w));
            }
        }
        // ---- end quirks of filter 2: out2[n2 - 10] = sum of the last 19 inputs, the rest stay 0
        {
            double o = 0;
            for (int k = 0; k < kBox; ++k) o += ring2[(r2 + k) % kBox][tx];
            pk.push(div_const<kBox>(o));
            for (int q = n2 - 9; q < n2; ++q) pk.push(0.0);
        }
        res.beat = pk.beat;
        // step 13 (reference src/tempo_atk_sort.c:283-287)
        const double tempo_score = (double)(4 * (float)pk.beat / (float)sd.duration) - 30.4;
        const double atk_score = -1.74 * atk_sum * 10000 / sd.n_samples + 58.3;
        res.tempo = (float)tempo_score;
        res.attack = (float)atk_score;
        } else if (p.what & BLX_DO_ENVELOPE) {
            res.tempo = nanf("");
            res.attack = nanf("");
        }
    }

    if (p.what == BLX_DO_ALL) { // reference src/analyze.c:68-79
        const float rating = (float)(fmax((double)res.tempo, 0.0) + (double)res.amplitude + (double)res.frequency +
                                     fmax((double)res.attack, 0.0));
        res.force = rating;
        res.calm_or_loud = (rating > 0) ? 0 : (rating < 0) ? 1 : 2;
    }
    if (in_range) p.out[s] = res;
}

// Step 6 of the envelope analyser for every hop of every song at once (reference
// src/tempo_atk_sort.c:186-190): x = log(1 + mu E) / log(1 + mu), mu = 100.0f. Element-wise over the
// packed energy rows; the sequential tail then only streams x.
// xlog carries 16 doubles of read-ahead slack behind the last row (the tail's prefetch); they are written as zeros.
__global__ void logcomp_kernel(const double *__restrict__ energy, double *__restrict__ xlog, long long n) {
    const float mu = 100.0f;
    const double log_den = log((double)(1 + mu));
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n + 16; i += (long long)gridDim.x * blockDim.x)
        xlog[i] = (i < n) ? log(1 + (double)mu * energy[i]) / log_den : 0.0;
}

cudaError_t launch_logcomp(const double *d_energy, double *d_xlog, long long n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int threads = 256;
    const long long blocks = (n + threads - 1) / threads;
    logcomp_kernel<<<(unsigned)(blocks > 148 * 32 ? 148 * 32 : blocks), threads, 0, st>>>(d_energy, d_xlog, n);
    return cudaGetLastError();
}

cudaError_t launch_tail(const TailParams &p, int n_songs, cudaStream_t st) {
    tail_kernel<<<(n_songs + kTailSongs - 1) / kTailSongs, kTailThreads, 0, st>>>(p, n_songs);
    return cudaGetLastError();
}

// =====================================================================================
// bl_rectangular_filter helper (sequential by definition: running sum)
// =====================================================================================
__global__ void rect_filter_kernel(double *out, const double *in, int n, int width) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
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

cudaError_t launch_rect_filter(double *d_out, const double *d_in, int n, int width, cudaStream_t st) {
    rect_filter_kernel<<<1, 32, 0, st>>>(d_out, d_in, n, width);
    return cudaGetLastError();
}

} // namespace blx
