// envelope.cu — hop energies of the envelope analyser (kernel id BLX_K_ENVELOPE).
//
// Replaces the hot loop of reference src/tempo_atk_sort.c:109-153: normalise the whole
// interleaved int16 stream to zero mean / unit variance, and for every hop of 256 samples
//   - run the 17-tap FIR (reference include/bandpass_coeffs.h:1-7) over a 512-sample window with
//     the delay line RESTARTED at the window start,
//   - take the 512-point double real FFT,
//   - sum |X_k|^2, k = 0..256, in a FLOAT accumulator in bin order (reference :142-149).
//
// Design. For j - window_start >= 16 the restarted FIR equals the continuous
// FIR of the stream (same operands), so the continuous FIR is computed ONCE per sample, in blocks
// of 256 samples. Every WARP works on its own: it owns a run of consecutive hops and walks it two
// hops (= two new FIR blocks) at a time; the block shared with the previous pair stays in a
// per-warp carry slot. The first 16 outputs of every window ("heads") are the continuous outputs
// minus the contribution of the 16 samples before the window, a 16 x 16 triangular product that
// the 16 lanes of the hop's half-warp evaluate from a coefficient table. Then the two FFTs of the
// pair run at once, 16 lanes each (fft16.cuh, double), followed by the float-accumulated power
// sum. The raw int16 samples of the next pair are fetched by the warp's own 1-D bulk copy
// (cp.async.bulk / UBLKCP, one mbarrier per buffer) while the FFTs of the current pair run. There
// is no block-wide barrier after set-up, so the warps of an SM drift apart and the FP64-bound
// phases (FFT) of some overlap the latency-bound phase (accumulation chain) of others.
//
// The FIR sums are EXACT: the reference's taps are decimal literals of seven digits, so c_k = C_k * 1e-7
// with integer C_k (|C_k| < 2^24), the samples are int16, and x = (s / 32768 - mean_d) / var_d (reference
// src/tempo_atk_sort.c:110-113) is affine in the raw sample s:
//   y = A * 1e-7 * (sum C_k s_k) - Bm * (sum c_k),  A = 1 / (32768 var_d), Bm = mean_d / var_d.
// sum C_k s_k < 2^44 is an integer whether it is accumulated in FP64 (the default: one DFMA per tap, no
// rounding below 2^53) or in 64-bit integers (BLX_ENV_FIR_INT), and it is normalised with one FMA per
// output. The result is the exact filter output rounded once; the reference rounds after every product and
// sum, so the two differ by ~1e-16 relative, like any two summation orders - nine orders of magnitude below
// the 1e-7 onset-count cliff.
//
// DUP = true: the stream is a mono signal u with every sample doubled (L = R), stored once
// (x[2 i] = x[2 i + 1] = u[i], the decimated stream pass 1 writes for float32 input). The 17 taps
// then fold into two 9-tap filters over u, one for even and one for odd outputs.
//
// The float accumulation s <- (float)((double)s + p_k), k = 0..256, is a chain of 257 dependent
// roundings (~50 cycles each on the FP64 + conversion pipes): run naively it idles the SM. It is
// evaluated by the 16 lanes that hold the hop's spectrum (float_chain below):
//   while s stays inside one binade [2^e, 2^(e+1)) its float grid is g = 2^(e-23) and
//   RN_g(s + p) = s + RN_g(p) because s is a multiple of g; so with q = s / g (a 24-bit integer)
//   the chain is q += I_k with I_k = RN_g(p_k) / g, which one double addition p_k + 1.5 * 2^52 * g
//   leaves in the low mantissa word.
// Which binade the sum is in when a bin is added is PREDICTED from the exact double prefix sum of the
// spectrum, in groups of four bins: a group whose prefix keeps its exponent is added as integers, the
// few others (~3 per hop) bin by bin with the reference's own step; every prediction is verified on
// the way and the rare hop that fails is redone by float_chain_rounds / chain_round, the first form
// of the algorithm (one prefix-scan round per binade).
// The result equals the reference's chain except when s + p_k lies within 2^-29 grid units of a
// rounding midpoint (the reference rounds to double, then to float; exact ties follow the parity of
// the magic constant instead of q).
#include <cstdio>
#include <cstdlib>
#include "blx_common.cuh"
#include "fft16.cuh"
#include "kernels.h"

namespace blx {

// Build-time switches for A/B measurements (tools/ab_envelope.sh); the defaults are the measured best.
#ifndef BLX_ENV_WARPS
#define BLX_ENV_WARPS 8 // warps per CTA, doubled-mono form
#endif
#ifndef BLX_ENV_WARPS_S16
#define BLX_ENV_WARPS_S16 8 // warps per CTA, interleaved stereo form (larger raw buffers: 8 is what two CTAs per SM hold)
#endif
#ifndef BLX_ENV_FIR_INT
// 0: the FIR sums in FP64 over the integer taps (DFMA, exact: every partial sum is an integer below 2^53);
// 1: in 64-bit integers (IMAD.WIDE). Measured on B200 (tools/ab_envelope.sh, 1 024 songs): 38.9 ms against 42.2 ms -
// ptxas splits every mad.wide into a multiply and a 3-input 64-bit add, two issue slots per tap, and the kernel
// is bound by dependent-issue latency, not by the FP64 pipe (40 % busy).
#define BLX_ENV_FIR_INT 0
#endif

#ifndef BLX_ENV_TW_SMEM
// 1: both twiddle tables (6 KB) are copied to shared memory by every CTA; 0: read through L1 (LDG)
#define BLX_ENV_TW_SMEM 1
#endif

#ifndef BLX_ENV_MINB
#define BLX_ENV_MINB 2 // CTAs per SM the register allocation is sized for (__launch_bounds__)
#endif
#ifndef BLX_ENV_TW_LOAD
// 1: the twiddles between the two FFT passes are applied behind the transpose (fft16.cuh, TW_ON_LOAD)
#define BLX_ENV_TW_LOAD 0
#endif
#ifndef BLX_ENV_TW2_FACTOR
// 1: W512^(l + 16 d) = W512^l * W32^d with W32^d as immediates: one table read per pair instead of eight, four more
// FP64 instructions per bin pair
#define BLX_ENV_TW2_FACTOR 0
#endif
#ifndef BLX_ENV_DIRTY_LANE0
// 1: the dirty groups of the accumulation are read by lane 0 of each half-warp only (two wavefronts per 128-bit read
// instead of four); the other lanes run the same instructions on zeros and their result is not used
#define BLX_ENV_DIRTY_LANE0 0
#endif
#ifndef BLX_ENV_PARTNER_SHFL
// 1: Z[256 - k] comes from its lane by shuffles (32 wavefronts) instead of through the exchange buffer (64)
#define BLX_ENV_PARTNER_SHFL 0
#endif

namespace {
#ifndef BLX_ENV_PAIRS
#define BLX_ENV_PAIRS 32
#endif
constexpr int kPairsPerWarp = BLX_ENV_PAIRS;     // a warp owns 2 * kPairsPerWarp consecutive hops (64)
constexpr int kSlotBytes = kHop * 8;             // one block of 256 FIR outputs (128 cells of 16 bytes)

// Shared-memory plan. Per warp: the FFT exchange buffers of its two hops (the first 2 KB first hold
// the FIR block B + 1 of the pair), the carry block, two raw-sample buffers and their mbarriers.
// A raw buffer holds r[i] = stream element blk * (B + 1) - pre + i, B = first block (= hop) of the pair:
// `pre` samples in front, then the two new blocks.
template <bool DUP> struct EG {
    static constexpr int warps = DUP ? BLX_ENV_WARPS : BLX_ENV_WARPS_S16; // independent warps per CTA
    static constexpr int threads = 32 * warps;
    static constexpr int hops_per_cta = 2 * kPairsPerWarp * warps;
    static constexpr int pre = DUP ? 8 : 16;     // raw elements in front of a block that its FIR reads
    static constexpr int blk = DUP ? 128 : 256;  // raw elements per block
    static constexpr int blk_bytes = blk * 2;
    static constexpr int per_thread = blk / 16;  // raw elements whose outputs one lane computes
    static constexpr int taps = DUP ? 8 : 16;    // head-correction terms per lane
    static constexpr int raw_bytes = ((pre + 2 * blk) * 2 + 63) / 64 * 64;
    static constexpr int w_xchg = 0;                                  // double2[2][272]; FIR block B + 1 at +0
    static constexpr int w_carry = w_xchg + 2 * kXchgElems * 16;      // FIR block B, then B + 2
    static constexpr int w_raw = w_carry + kSlotBytes;                // 2 raw buffers
    static constexpr int w_bar = w_raw + 2 * raw_bytes;               // 2 mbarriers
    static constexpr int w_bytes = (w_bar + 16 + 127) / 128 * 128;
    static constexpr int off_tabi = warps * w_bytes;              // int[16][taps] head table (integer taps)
    static constexpr int off_tabd = off_tabi + 16 * taps * 4;         // double[16]: sum of the dropped taps
    static constexpr int off_tw1 = off_tabd + 16 * 8;                 // double2[256], double2[128] (BLX_ENV_TW_SMEM)
    static constexpr int off_tw2 = off_tw1 + 256 * 16;
    static constexpr int bytes = BLX_ENV_TW_SMEM ? off_tw2 + 128 * 16 : off_tw1;
    static_assert(bytes <= 115712, "two CTAs per SM");
};

// reference include/bandpass_coeffs.h:1-7 — coeffs[0][0..8]; the filter is symmetric. The literals have seven
// decimals: kTapI[k] = coeffs[0][k] * 1e7 exactly.
__device__ __forceinline__ constexpr int tap_i(int k) {
    constexpr int c[9] = {-23470, 44613, -114627, 226382, -405147, 580037, -779167, 882711, 9065095};
    return c[k];
}
__constant__ double c_fir_half[9] = {-0.0023470, 0.0044613, -0.0114627, 0.0226382, -0.0405147,
                                     0.0580037,  -0.0779167, 0.0882711, 0.9065095};
__constant__ int c_fir_half_i[9] = {-23470, 44613, -114627, 226382, -405147, 580037, -779167, 882711, 9065095};
// Integer coefficient that multiplies x[j - m] in the 17-tap FIR, m = 0..16 (symmetric; reference
// include/bandpass_coeffs.h:1-7 and the loop of reference src/tempo_atk_sort.c:124-137).
__device__ __forceinline__ constexpr int lag_i(int m) { return tap_i(m <= 8 ? m : 16 - m); }
// Folded taps of the doubled mono stream: even output 2 n = sum_m fold_e(m) u[n - m], odd output
// 2 n + 1 = sum_m fold_d(m) u[n - m], m = 0..8.
__device__ __forceinline__ constexpr int fold_e_i(int m) { return m == 0 ? lag_i(0) : lag_i(2 * m - 1) + lag_i(2 * m); }
__device__ __forceinline__ constexpr int fold_d_i(int m) { return lag_i(2 * m) + (m < 8 ? lag_i(2 * m + 1) : 0); }
// the same with a run-time argument (table set-up only), integer and double
__device__ __forceinline__ int lag_i_rt(int m) { return c_fir_half_i[m <= 8 ? m : 16 - m]; }
__device__ __forceinline__ int fold_e_i_rt(int m) { return m == 0 ? lag_i_rt(0) : lag_i_rt(2 * m - 1) + lag_i_rt(2 * m); }
__device__ __forceinline__ int fold_d_i_rt(int m) { return lag_i_rt(2 * m) + (m < 8 ? lag_i_rt(2 * m + 1) : 0); }
__device__ __forceinline__ double lag_d_rt(int m) { return c_fir_half[m <= 8 ? m : 16 - m]; }
__device__ __forceinline__ double fold_e_d_rt(int m) { return m == 0 ? lag_d_rt(0) : lag_d_rt(2 * m - 1) + lag_d_rt(2 * m); }
__device__ __forceinline__ double fold_d_d_rt(int m) { return lag_d_rt(2 * m) + (m < 8 ? lag_d_rt(2 * m + 1) : 0.0); }

// acc + x * C / acc + x * c as ONE 32 x 32 + 64-bit multiply-add (IMAD.WIDE); the C++ form (long long)c * x widens first
// and costs a 64-bit multiply sequence
template <int C> __device__ __forceinline__ long long madw(int x, long long acc) {
    asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x), "n"(C));
    return acc;
}
__device__ __forceinline__ long long madw_r(int x, int c, long long acc) {
    asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x), "r"(c));
    return acc;
}
// the nine folded taps of output pair `o` of a lane (doubled mono stream), m = M..8
template <int M> __device__ __forceinline__ void fir_dup_taps(const int (&u)[16], int o, long long &ye, long long &yd) {
    ye = madw<fold_e_i(M)>(u[8 + o - M], ye);
    yd = madw<fold_d_i(M)>(u[8 + o - M], yd);
    if constexpr (M < 8) fir_dup_taps<M + 1>(u, o, ye, yd);
}
// the symmetric tap pairs k = K..7 of output `o` of a lane (interleaved stereo stream)
template <int K> __device__ __forceinline__ void fir_sym_taps(const int (&xv)[32], int o, long long &y) {
    y = madw<tap_i(K)>(xv[o + 16 - K] + xv[o + K], y);
    if constexpr (K < 7) fir_sym_taps<K + 1>(xv, o, y);
}

// (double)v for |v| < 2^51: one integer add on the high word (the constant's low word is 0) and one DADD.
__device__ __forceinline__ double i64_to_double(long long v) {
    return __longlong_as_double(v + 0x4338000000000000ll) - 6755399441055744.0;
}

// ---------------------------------------------------------------- power spectrum layout
// Where bin k (0..256) of a hop's power spectrum lives in its exchange buffer (doubles). Lane b of the
// accumulation owns bins 16 b + 1 .. 16 b + 16 = row b, 18 doubles apart (16-byte aligned rows, conflict-free
// 128-bit reads): element i < 15 at slot i, element 15 at slot 17 in rows 0..6 and at slot 15 in rows 7..15
// (whichever keeps the 16 lanes that write one bin each - bins l + 16 d, or 256 - l - 16 d - on 16 different
// banks); bin 0 at slot 15 of row 0.
constexpr int kPRow = 18;
__host__ __device__ constexpr int pslot15(int b) { return b <= 6 ? 17 : 15; }
__device__ __forceinline__ int pidx(int k) {
    if (k == 0) return 15;
    const int b = (k - 1) >> 4, i = (k - 1) & 15;
    return kPRow * b + (i == 15 ? pslot15(b) : i);
}
static_assert(kPRow * 16 * 8 <= kXchgElems * 16, "spectrum fits the exchange buffer");

// This lane's 16 bins (row lane16) from shared memory: nine 128-bit loads.
__device__ __forceinline__ void load_row(const double *xr, int lane16, double (&pv)[16]) {
    const double2 *row = reinterpret_cast<const double2 *>(xr + kPRow * lane16);
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const double2 t = row[i];
        pv[2 * i] = t.x;
        pv[2 * i + 1] = t.y;
    }
    const double2 a = row[7], b = row[8];
    pv[14] = a.x;
    pv[15] = (lane16 <= 6) ? b.y : a.y;
}

// sum_fft of reference src/tempo_atk_sort.c:142-150 for the two hops of a warp, one per half-warp (see
// the header comment). xr[pidx(k)] = |X_k|^2 of this lane's half-warp; `active` is false for a
// half-warp without a hop (last tile of a song). Returns (double)sum_fft on every lane of the half.
// All 32 lanes run the same control flow, so every shuffle / vote uses the full mask (a partial-mask
// shuffle compiles to a divergence-safe sequence several times slower); the state of a chain is
// replicated in the 16 lanes of its half-warp. One loop iteration = one binade:
//   1. every lane adds up the increments of its own 16 bins (lanes before the current position count
//      nothing; the one partly consumed lane subtracts what the half-warp measures for its consumed bins),
//   2. a prefix scan over the 16 lane sums finds the first lane that takes the sum out of the binade,
//   3. the 16 lanes take one bin of that lane each to find the bin k* itself,
//   4. bin k* is added with the reference's double-add / float-convert step, which gives the next binade.
template <bool GUARD>
__device__ __forceinline__ void chain_round(const double *xr, const double (&pv)[16], int lane16, int lane_base, double &r,
                                            int &kdone, bool &done) {
    const unsigned full = 0xffffffffu;
    const int rhi = __double2hiint(r), rlo = __double2loint(r);
    const int ex = (rhi >> 20) & 0x7ff;
    const bool normal = (ex >= 1023 - 126) && (ex <= 1023 + 127);
    if (!done && !normal) {
        // zero, subnormal, inf or nan: no binade to scan in; one plain reference step
        r = (double)(float)(r + xr[pidx(kdone + 1)]);
        kdone += 1;
        if (kdone == 256) done = true;
    }
    const bool scan = !done && normal;
    // binade e = ex - 1023, float grid g = 2^(e-23); r = q g with 2^23 <= q < 2^24
    const int hiM = ((ex + 29) << 20) | 0x80000; // M = 1.5 * 2^(e+29): ulp(M) = g
    const double M = __hiloint2double(hiM, 0);
    const int q = ((rhi & 0xFFFFF) << 3) | (int)((unsigned)rlo >> 29) | 0x800000;
    auto increment = [&](double p) -> int { // RN_g(p) / g (ties follow M's parity)
        const double t = p + M;
        unsigned v = (unsigned)__double2loint(t);
        if (GUARD) { // a bin of 2^25 grid units or more: cap it (it crosses anyway)
            if (__double2hiint(t) != hiM) v = 1u << 25;
            v = min(v, 1u << 25);
        }
        return (int)v;
    };
    // (1) lane sums
    int lane_sum = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) lane_sum += increment(pv[i]);
    if (GUARD) lane_sum = min(lane_sum, 1 << 25);
    const int cur = (kdone - 1) >> 4; // lane that holds bin kdone; its bins up to kdone are consumed
    {
        const int kk = 16 * cur + 1 + lane16; // the half-warp measures the consumed part of lane `cur`
        const int c = (kk <= kdone) ? increment(xr[pidx(kk)]) : 0;
        const int c_lo = __reduce_add_sync(full, (threadIdx.x & 16) ? 0 : c);
        const int c_hi = __reduce_add_sync(full, (threadIdx.x & 16) ? c : 0);
        const int consumed = (threadIdx.x & 16) ? c_hi : c_lo;
        if (lane16 == cur) lane_sum = GUARD ? max(lane_sum - consumed, 0) : lane_sum - consumed;
        if (lane16 < cur || !scan) lane_sum = 0;
    }
    // (2) first lane that leaves the binade
    int incl = lane_sum;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int up = __shfl_up_sync(full, incl, o, 16);
        if (lane16 >= o) incl += up;
    }
    const unsigned lanes_over = (__ballot_sync(full, scan && q + incl >= (1 << 24)) >> lane_base) & 0xFFFFu;
    const int src = lanes_over ? (__ffs(lanes_over) - 1) : 15;
    const int base = __shfl_sync(full, q + incl - (lanes_over ? lane_sum : 0), src, 16); // sum before lane src
    if (!__any_sync(full, lanes_over != 0u)) { // both chains of the warp end inside their binades
        if (scan) {
            r = __hiloint2double((ex << 20) | ((base & 0x7FFFFF) >> 3), (base & 7) << 29);
            done = true;
        }
        return;
    }
    // (3) the bin inside lane src
    const int kk = 16 * src + 1 + lane16;
    const int inc1 = (scan && lanes_over && kk > kdone) ? increment(xr[pidx(kk)]) : 0;
    int incl1 = inc1;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int up = __shfl_up_sync(full, incl1, o, 16);
        if (lane16 >= o) incl1 += up;
    }
    const unsigned bins_over = (__ballot_sync(full, scan && lanes_over && base + incl1 >= (1 << 24)) >> lane_base) & 0xFFFFu;
    const int bsrc = bins_over ? (__ffs(bins_over) - 1) : 15;
    const int qb = __shfl_sync(full, base + incl1 - inc1, bsrc, 16); // sum before bin k*
    // (4)
    if (scan) {
        if (lanes_over) {
            const int kstar = 16 * src + 1 + bsrc;
            const double sq = __hiloint2double((ex << 20) | ((qb & 0x7FFFFF) >> 3), (qb & 7) << 29);
            r = (double)(float)(sq + xr[pidx(kstar)]); // reference src/tempo_atk_sort.c:147
            kdone = kstar;
            if (kdone == 256) done = true;
        } else {
            r = __hiloint2double((ex << 20) | ((base & 0x7FFFFF) >> 3), (base & 7) << 29);
            done = true;
        }
    }
}

// Slow path of the accumulation (rare, see float_chain): the binade-by-binade scan from bin 17 on,
// r16 = the sum after bin 16. Called by all 32 lanes.
__device__ __noinline__ double float_chain_rounds(const double *xr, int lane16, bool active, double r16) {
    const unsigned full = 0xffffffffu;
    const int lane_base = (threadIdx.x & 16);
    double pv[16];
    load_row(xr, lane16, pv); // bins 16 lane16 + 1 + i
    double r = r16;
    int kdone = 16; // bins 0..kdone are in r
    bool done = !active;
    while (__any_sync(full, !done)) chain_round<true>(xr, pv, lane16, lane_base, r, kdone, done);
    return r;
}

// the float q (an integer 2^23 <= q < 2^24 on the grid of binade `ex`) as a double
__device__ __forceinline__ double grid_to_double(int ex, unsigned q) {
    return __hiloint2double((ex << 20) | (int)((q & 0x7FFFFFu) >> 3), (int)((q & 7u) << 29));
}

// sum_fft of reference src/tempo_atk_sort.c:142-150 for the two hops of a warp, one per half-warp:
// s <- (float)((double)s + p_k), k = 0..256, with p_k = xr[pidx(k)]. `active` is false for a half-warp
// without a hop. Returns (double)sum_fft on every lane of the half-warp.
//
// While s stays inside one binade [2^e, 2^(e+1)) its float grid is g = 2^(e-23) and RN_g(s + p) =
// s + RN_g(p), so with q = s / g the chain is the integer sum q += I_k, I_k = RN_g(p_k) / g, which one
// double addition p_k + 1.5 * 2^52 * g leaves in the low mantissa word. Which binade s is in when a bin
// is added is PREDICTED from the exact prefix sum S of the p_k in double (the float chain stays within
// 257 * 2^-25 relative of it): lane b owns bins 16 b + 1 .. 16 b + 16 in four groups of four, a double
// prefix scan over the half-warp gives S at every group boundary. A group over which S keeps its exponent
// is CLEAN: its bins are converted at that binade's grid and added as integers (an integer prefix scan
// gives the sum of all clean bins in front of any group). The other groups (~3 per hop) are DIRTY: their
// four bins are added one after the other, in order, with the reference's own double-add / float-convert
// step. Bins 0..16 (lane 0), where the sum climbs a binade per bin, run as the plain chain meanwhile.
// Every prediction is VERIFIED on the way: in front of a dirty group the integer sum is below 2^24 and
// the chain is in the predicted binade, behind it the chain is in the binade predicted for the next
// group, and the same at the end; clean neighbours share their boundary exponent by construction. A hop
// that fails (a prefix sum within rounding noise of a power of two; sums that are not normal floats) is
// redone by the binade-by-binade scan above. Host model and test against the sequential chain:
// tools/chain_model4.c.
#ifndef BLX_ENV_CVT_MAGIC
#define BLX_ENV_CVT_MAGIC 0 // 1: int -> double by a magic-constant add (ALU + FP64 pipes) instead of I2F.F64 (conversion pipe)
#endif
#ifndef BLX_ENV_CHAIN_G
#define BLX_ENV_CHAIN_G 4 // bins per group of the accumulation (4 or 2)
#endif
// scratch behind the spectrum rows of a hop's exchange buffer: per lane and group {clean bins in front, exponents}
constexpr int kAuxOff = kPRow * 16;          // doubles
constexpr int kAuxLaneStride = 10;           // doubles per lane (8 used for G = 2): 80-byte rows, conflict-free 128-bit stores
static_assert((kAuxOff + 16 * kAuxLaneStride) * 8 <= kXchgElems * 16, "scratch fits behind the spectrum");

template <int G>
__device__ __forceinline__ double float_chain(double *xr, int lane16, bool active, bool force_slow) {
    constexpr int NG = 16 / G;                     // groups per lane
    constexpr int kWords = 16 * NG / 32;           // 32-bit words of one half-warp's dirty mask
    constexpr int kLanesPerWord = 32 / NG;
    const unsigned full = 0xffffffffu;
    const int hw = (threadIdx.x >> 4) & 1;
    double pv[16];
    load_row(xr, lane16, pv); // bins 16 lane16 + 1 + i
    const double p0 = xr[15]; // bin 0
    // bins 0..16 one by one on lane 0's registers (the other lanes run the same instructions on their own bins;
    // only lane 0's result is used). Independent of everything up to the resolution below.
    double r16;
    {
        float sf = (float)p0;
#pragma unroll
        for (int k = 0; k < 16; ++k) sf = (float)((double)sf + pv[k]);
        r16 = (double)sf;
    }
    r16 = __shfl_sync(full, r16, 0, 16);
    // exact prefix sums in double: local (kept at the group boundaries), then across the half-warp
    double cb[NG], inc;
    {
        double c = pv[0] + (lane16 == 0 ? p0 : 0.0);
#pragma unroll
        for (int i = 1; i < 16; ++i) {
            if (i % G == 0) cb[i / G - 1] = c;
            c += pv[i];
        }
        cb[NG - 1] = c;
        inc = c;
    }
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const double up = __shfl_up_sync(full, inc, o, 16);
        if (lane16 >= o) inc += up;
    }
    double excl = __shfl_up_sync(full, inc, 1, 16);
    if (lane16 == 0) excl = 0.0;
    // Nothing left to add after bin 16 (digital silence, or bins below half a double ulp of the sum): the
    // chain ends at r16. (Without this a silent hop, whose sum never becomes a normal float, would take the
    // slow path one bin at a time.)
    const bool rest_zero = __shfl_sync(full, inc, 15, 16) == __shfl_sync(full, inc, 0, 16);
    // the groups: exponent of S at the boundaries, clean groups summed at their grid
    const bool mine = active && lane16 != 0;
    int h[NG + 1];
    h[0] = __double2hiint(excl);
#pragma unroll
    for (int j = 0; j < NG - 1; ++j) h[j + 1] = __double2hiint(excl + cb[j]);
    h[NG] = __double2hiint(inc);
    unsigned before[NG], dirty = 0, run = 0;
#pragma unroll
    for (int j = 0; j < NG; ++j) {
        const double M = __hiloint2double((h[j] & 0x7FF00000) + 0x1D80000, 0); // 1.5 * 2^(e + 29): ulp = float grid of binade e
        unsigned g = 0;
#pragma unroll
        for (int i = G * j; i < G * j + G; ++i) g += (unsigned)__double2loint(pv[i] + M);
        const bool d = mine && (((h[j] ^ h[j + 1]) & 0x7FF00000) != 0);
        before[j] = run;
        run += (d || !mine) ? 0u : g;
        dirty |= d ? (1u << j) : 0u;
    }
    unsigned incI = run;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const unsigned up = __shfl_up_sync(full, incI, o, 16);
        if (lane16 >= o) incI += up;
    }
    const unsigned baseI = incI - run;
    const unsigned total = __shfl_sync(full, incI, 15, 16);
    const int e_end = __shfl_sync(full, h[NG], 15, 16) >> 20;
    // per lane and group: clean bins in front of the group, exponents in front of / behind it (read back by the
    // half-warp for the dirty groups only)
    {
        uint2 *aux = reinterpret_cast<uint2 *>(xr + kAuxOff + kAuxLaneStride * lane16);
#pragma unroll
        for (int j = 0; j < NG; j += 2) {
            const uint4 t = make_uint4(baseI + before[j], ((unsigned)h[j] >> 20) | (((unsigned)h[j + 1] >> 20) << 16),
                                       baseI + before[j + 1], ((unsigned)h[j + 1] >> 20) | (((unsigned)h[j + 2] >> 20) << 16));
            *reinterpret_cast<uint4 *>(aux + j) = t;
        }
    }
    // the dirty groups of this half-warp, bit NG b + j, as kWords words
    unsigned mk[kWords];
    {
        const unsigned bits = dirty << (NG * (lane16 % kLanesPerWord));
        const int word = kWords * hw + lane16 / kLanesPerWord;
#pragma unroll
        for (int w = 0; w < kWords; ++w) {
            const unsigned lo = __reduce_or_sync(full, word == w ? bits : 0u), hi = __reduce_or_sync(full, word == kWords + w ? bits : 0u);
            mk[w] = hw ? hi : lo;
        }
    }
    __syncwarp(full); // scratch visible
    // the dirty groups in order; both half-warps step together (the one that runs out idles)
    int ex = (__double2hiint(r16) >> 20) & 0x7ff;
    unsigned q = (((unsigned)__double2hiint(r16) & 0xFFFFFu) << 3) | ((unsigned)__double2loint(r16) >> 29) | 0x800000u;
    bool ok = !active || (ex >= 1023 - 126 && ex <= 1023 + 126);
    unsigned Pprev = 0;
    for (;;) {
        unsigned any = mk[0];
#pragma unroll
        for (int w = 1; w < kWords; ++w) any |= mk[w];
        if (!__any_sync(full, any != 0u)) break;
        const bool act = any != 0u;
        int bit = NG; // idle: lane 1, group 0
        bool found = false;
#pragma unroll
        for (int w = 0; w < kWords; ++w) {
            const bool take = !found && mk[w] != 0u;
            if (take) { bit = 32 * w + (__ffs(mk[w]) - 1); mk[w] &= mk[w] - 1; }
            found = found || take;
        }
        const int b = bit / NG, j = bit % NG;
        const bool reader = !BLX_ENV_DIRTY_LANE0 || lane16 == 0;
        uint2 rec = make_uint2(0u, 0u);
        if (reader) rec = *reinterpret_cast<const uint2 *>(xr + kAuxOff + kAuxLaneStride * b + j);
        const unsigned Pb = rec.x, pk = rec.y;
        // the group's bins (row b, elements G j .. G j + G - 1)
        const double2 *row = reinterpret_cast<const double2 *>(xr + kPRow * b);
        double pg[G];
        {
            double2 pa = make_double2(0.0, 0.0), pb = pa, pc = pa;
            if (reader) {
                pa = row[(G / 2) * j];
                pc = row[8];
                if (G == 4) pb = row[2 * j + 1];
            }
            pg[0] = pa.x; pg[1] = pa.y;
            if (G == 4) { pg[G - 2] = pb.x; pg[G - 1] = pb.y; }
            if (j == NG - 1 && b <= 6) pg[G - 1] = pc.y;
        }
        if (act) {
            const unsigned qb = q + (Pb - Pprev);
            ok = ok && qb < (1u << 24) && (int)(pk & 0x7FFu) == ex;
            double r = grid_to_double(ex, qb);
#pragma unroll
            for (int i = 0; i < G; ++i) r = (double)(float)(r + pg[i]); // reference src/tempo_atk_sort.c:147
            ex = (__double2hiint(r) >> 20) & 0x7ff;
            q = (((unsigned)__double2hiint(r) & 0xFFFFFu) << 3) | ((unsigned)__double2loint(r) >> 29) | 0x800000u;
            ok = ok && (int)((pk >> 16) & 0x7FFu) == ex && ex <= 1023 + 126;
            Pprev = Pb;
        }
    }
    const unsigned qf = q + (total - Pprev);
    ok = ok && (!active || (qf < (1u << 24) && e_end == ex && ex <= 1023 + 126));
    double r = grid_to_double(ex, qf);
    if (rest_zero) { r = r16; ok = true; }
    if (BLX_ENV_DIRTY_LANE0 && lane16 != 0) ok = true; // only lane 0 of a half-warp followed the dirty groups
    if (__any_sync(full, !ok) || force_slow) r = float_chain_rounds(xr, lane16, active, r16);
    return r;
}

} // namespace

#ifdef BLX_ENV_MAXNREG // A/B only: an explicit register cap instead of the one __launch_bounds__ derives
#define BLX_ENV_BOUNDS __maxnreg__(BLX_ENV_MAXNREG)
#else
#define BLX_ENV_BOUNDS __launch_bounds__(EG<DUP>::threads, BLX_ENV_MINB)
#endif
template <bool DUP> __global__ void BLX_ENV_BOUNDS envelope_kernel(EnvelopeParams p) {
    using G = EG<DUP>;
    extern __shared__ __align__(128) unsigned char smem[];
    const SongDesc sd = p.songs[blockIdx.y];
    if (blockIdx.x * G::hops_per_cta >= sd.n_hops) return;
    const SongNorm nm = p.norm[blockIdx.y];
    if (nm.status != 0) return;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int hw = lane >> 4, lane16 = lane & 15; // half-warp = hop of the pair
    int *tabi = reinterpret_cast<int *>(smem + G::off_tabi);
    double *tabd = reinterpret_cast<double *>(smem + G::off_tabd);
    // Head table: a window's output t (its first 16) lacks the terms of the `pre` raw samples X[0..pre)
    // in front of the window; row `lane16` holds the integer coefficients of those terms for the output this
    // lane corrects, tabd the sum of the dropped coefficients (for the mean term).
    //   DUP: lane = n + 8 part, output 2 n + part: sum_{j >= n} fold(8 + n - j) X[j]
    //   else: lane = output j:                      sum_{i >= j} c(j + 16 - i) X[i]
    if (tid < 16) {
        int *trow = tabi + tid * G::taps;
        double dsum = 0.0;
        if (DUP) {
            const int n = tid & 7, part = tid >> 3;
            for (int j = 0; j < 8; ++j) {
                const int m = 8 + n - j;
                trow[j] = (j >= n) ? (part ? fold_d_i_rt(m) : fold_e_i_rt(m)) : 0;
                dsum += (j >= n) ? (part ? fold_d_d_rt(m) : fold_e_d_rt(m)) : 0.0;
            }
        } else {
            for (int i = 0; i < 16; ++i) {
                trow[i] = (i >= tid) ? lag_i_rt(tid + 16 - i) : 0;
                dsum += (i >= tid) ? lag_d_rt(tid + 16 - i) : 0.0;
            }
        }
        tabd[tid] = dsum;
    }
#if BLX_ENV_TW_SMEM
    {
        double2 *t1 = reinterpret_cast<double2 *>(smem + G::off_tw1);
        for (int i = tid; i < 256 + 128; i += G::threads) t1[i] = i < 256 ? p.tw1[i] : p.tw2[i - 256];
    }
    const double2 *tw1 = reinterpret_cast<const double2 *>(smem + G::off_tw1);
    const double2 *tw2 = reinterpret_cast<const double2 *>(smem + G::off_tw2);
#else
    const double2 *tw1 = p.tw1, *tw2 = p.tw2;
#endif

    unsigned char *wsm = smem + warp * G::w_bytes;
    double2 *xchg = reinterpret_cast<double2 *>(wsm + G::w_xchg) + hw * kXchgElems;
    unsigned char *newblk = wsm + G::w_xchg; // FIR block B + 1 of the pair (consumed before the FFT exchange)
    unsigned char *carry = wsm + G::w_carry; // FIR block B, overwritten with block B + 2 once hop 0 holds B in registers
    unsigned char *raw = wsm + G::w_raw;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wsm + G::w_bar);

    const int B0 = blockIdx.x * G::hops_per_cta + warp * (2 * kPairsPerWarp); // first hop (= first block) of this warp
    const int n_mine = min(2 * kPairsPerWarp, sd.n_hops - B0);            // hops of this warp
    const int n_pairs = (n_mine + 1) >> 1;
    const short *stream = p.stream + (DUP ? sd.q_off : sd.pcm_off);
    // Pass q (lane 0): q >= 0 stages the raw samples of the new blocks of pair q; the prologue q = -1 stages
    // block B0 where the upper half-warp expects it. Buffer = q & 1.
    auto issue_pass = [&](int q) {
        unsigned char *dst = raw + (q & 1) * G::raw_bytes;
        long long e0;
        int bytes;
        if (q < 0) {
            e0 = (long long)G::blk * B0 - G::pre;
            dst += G::blk_bytes;
            bytes = (G::blk + G::pre) * 2;
            if (B0 == 0) { e0 = 0; dst += G::pre * 2; bytes = G::blk * 2; } // nothing before the song: zeros (set below)
        } else {
            const int nb = min(2, n_mine - 2 * q); // new blocks this pair needs; (m + 2) * 256 <= 512 F <= n
            e0 = (long long)G::blk * (B0 + 2 * q + 1) - G::pre;
            bytes = (nb * G::blk + G::pre) * 2;
        }
        fence_proxy_async();
        mbar_arrive_expect_tx(bar + (q & 1), (unsigned)bytes);
        tma_load_1d(dst, stream + e0, (unsigned)bytes, bar + (q & 1));
    };
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_fence_init();
        if (n_mine > 0) {
            if (B0 == 0) {
                int4 *z = reinterpret_cast<int4 *>(raw + G::raw_bytes + G::blk_bytes);
                z[0] = make_int4(0, 0, 0, 0);
                if (!DUP) z[1] = make_int4(0, 0, 0, 0);
            }
            issue_pass(-1);
            issue_pass(0);
        }
    }
    __syncthreads(); // the only block-wide barrier: table and mbarriers are set up
    if (n_mine <= 0) return;

    // y = A7 * (sum C_k s_k) - Bm * (sum c_k) (header comment). Both carry an extra factor 1/2 (exact): it is
    // the 1/2 of the real-FFT even/odd split, so the split below produces X_k without its 0.25 |.|^2 scaling
    // step; the three purely real bins are scaled back.
    const double A7 = nm.inv_var_d * (1.0 / 32768) * 0.5 * 1e-7;
    const double Bm = nm.mean_d * nm.inv_var_d * 0.5;
    double csum_all = c_fir_half[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) csum_all += 2.0 * c_fir_half[k];
    const double Ball = Bm * csum_all;
    const unsigned full = 0xffffffffu;
    const int key = 16 * (lane16 & 7);
    // byte offset of ring cell 16 a + lane16 inside a block: 256 a + 16 (lane16 ^ ((2 a + (lane16 >> 3)) & 7))
    int coff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) coff[i] = 16 * (lane16 ^ ((2 * i + (lane16 >> 3)) & 7));
    // the head correction's own constants
    const double head_mean = Bm * tabd[lane16];

    for (int q = -1; q < n_pairs; ++q) {
        const unsigned char *rcur = raw + (q & 1) * G::raw_bytes;
        double2 v[16];
        // ---- hop 0 takes block B out of the carry slot before the upper half-warp overwrites it with B + 2
        if (q >= 0 && hw == 0) {
#pragma unroll
            for (int a = 0; a < 8; ++a) v[a] = *reinterpret_cast<const double2 *>(carry + 256 * a + coff[a & 3]);
        }
        mbar_wait(bar + (q & 1), (unsigned)((q + 1) >> 1) & 1u);
        __syncwarp(full);

        // ---- continuous FIR: half-warp hw computes block B + 1 + hw, every lane 16 consecutive outputs
        // (8 cells); in the prologue only the upper half-warp's block (B0) is kept
        {
            const unsigned char *sp = rcur + lane * (G::per_thread * 2); // r[per_thread * lane ...]
            unsigned char *slot = (hw ? carry : newblk) + 128 * lane16;
            const bool keep = q >= 0 || hw == 1;
            if (DUP) {
                // inputs u[k] = r[8 lane + k], k = 0..15; output pair o uses u[8 + o - m], m = 0..8
                const int4 u0 = reinterpret_cast<const int4 *>(sp)[0], u1 = reinterpret_cast<const int4 *>(sp)[1];
                const int wds[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
                int u[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    u[2 * i] = (int)(short)(wds[i] & 0xffff);
                    u[2 * i + 1] = wds[i] >> 16;
                }
#if BLX_ENV_FIR_INT
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    long long ye = 0, yd = 0;
                    fir_dup_taps<0>(u, o, ye, yd);
                    const double2 y2 = make_double2(fma(i64_to_double(ye), A7, -Ball), fma(i64_to_double(yd), A7, -Ball));
                    if (keep) *reinterpret_cast<double2 *>(slot + ((16 * o) ^ key)) = y2;
                }
#else
                double ud[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) ud[i] = BLX_ENV_CVT_MAGIC ? int_to_double_exact(u[i]) : (double)u[i];
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    double ye = (double)fold_e_i(0) * ud[8 + o], yd = (double)fold_d_i(0) * ud[8 + o];
#pragma unroll
                    for (int m = 1; m <= 8; ++m) {
                        ye = fma((double)fold_e_i(m), ud[8 + o - m], ye);
                        yd = fma((double)fold_d_i(m), ud[8 + o - m], yd);
                    }
                    const double2 y2 = make_double2(fma(ye, A7, -Ball), fma(yd, A7, -Ball));
                    if (keep) *reinterpret_cast<double2 *>(slot + ((16 * o) ^ key)) = y2;
                }
#endif
            } else {
                // inputs xv[k] = r[16 lane + k], k = 0..31; output o uses xv[o + 16 - m], m = 0..16
                // (reference src/tempo_atk_sort.c:124-137; exact, so the order of the sum is free)
                int xv[32];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int4 q4 = reinterpret_cast<const int4 *>(sp)[i];
                    const int wds[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        xv[8 * i + 2 * j] = (int)(short)(wds[j] & 0xffff);
                        xv[8 * i + 2 * j + 1] = wds[j] >> 16;
                    }
                }
#if !BLX_ENV_FIR_INT
                double xd[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) xd[i] = BLX_ENV_CVT_MAGIC ? int_to_double_exact(xv[i]) : (double)xv[i];
#endif
#pragma unroll
                for (int o2 = 0; o2 < 8; ++o2) {
                    double yy[2];
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const int o = 2 * o2 + t;
#if BLX_ENV_FIR_INT
                        long long y = madw<tap_i(8)>(xv[o + 8], 0ll);
                        fir_sym_taps<0>(xv, o, y);
                        yy[t] = fma(i64_to_double(y), A7, -Ball);
#else
                        double y = (double)tap_i(8) * xd[o + 8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) y = fma((double)tap_i(k), xd[o + 16 - k] + xd[o + k], y);
                        yy[t] = fma(y, A7, -Ball);
#endif
                    }
                    if (keep) *reinterpret_cast<double2 *>(slot + ((16 * o2) ^ key)) = make_double2(yy[0], yy[1]);
                }
            }
        }
        __syncwarp(full);
        if (q < 0) continue;

        // ---- FFT input: the 512 FIR outputs of this half-warp's window: blocks B, B + 1 (hop 0, B already in
        // registers) or B + 1, B + 2 (hop 1) ...
        {
            const unsigned char *upper = hw ? carry : newblk; // second block of the window: B + 2 / B + 1
#pragma unroll
            for (int a = 0; a < 8; ++a) v[a + 8] = *reinterpret_cast<const double2 *>(upper + 256 * a + coff[a & 3]);
            if (hw) {
#pragma unroll
                for (int a = 0; a < 8; ++a) v[a] = *reinterpret_cast<const double2 *>(newblk + 256 * a + coff[a & 3]);
            }
        }
        // ... whose first 16 see an empty delay line: take out the terms of the raw samples in front of
        // the window (head table above) and put back the mean term of the dropped coefficients. Hop 1's
        // window starts at r[pre]; hop 0's one block earlier, at r'[blk + pre] of the previous pass.
        {
            const unsigned char *hp = hw ? rcur : raw + ((q & 1) ^ 1) * G::raw_bytes + G::blk_bytes;
            const int4 *trow = reinterpret_cast<const int4 *>(tabi + lane16 * G::taps);
            long long corr = 0;
#pragma unroll
            for (int i = 0; i < G::taps / 8; ++i) {
                const int4 q4 = reinterpret_cast<const int4 *>(hp)[i];
                const int4 ca = trow[2 * i], cb = trow[2 * i + 1];
                corr = madw_r((int)(short)(q4.x & 0xffff), ca.x, corr);
                corr = madw_r(q4.x >> 16, ca.y, corr);
                corr = madw_r((int)(short)(q4.y & 0xffff), ca.z, corr);
                corr = madw_r(q4.y >> 16, ca.w, corr);
                corr = madw_r((int)(short)(q4.z & 0xffff), cb.x, corr);
                corr = madw_r(q4.z >> 16, cb.y, corr);
                corr = madw_r((int)(short)(q4.w & 0xffff), cb.z, corr);
                corr = madw_r(q4.w >> 16, cb.w, corr);
            }
            const double h = fma(-A7, i64_to_double(corr), head_mean);
            const int s0 = DUP ? lane16 : 2 * lane16, s1 = DUP ? lane16 + 8 : 2 * lane16 + 1;
            const double h0 = __shfl_sync(full, h, s0 & 15, 16);
            const double h1 = __shfl_sync(full, h, s1 & 15, 16);
            if (lane16 < 8) { v[0].x += h0; v[0].y += h1; }
        }
        __syncwarp(full); // FIR blocks and raw samples are consumed
        if (lane == 0 && q + 1 < n_pairs) issue_pass(q + 1);

        // ---- 2 x (512-point double real FFT + float-accumulated power), one hop per half-warp in
        // lockstep (a half-warp without a hop works on stale data and is ignored)
        {
            const int hop = B0 + 2 * q + hw;
            const bool active = 2 * q + hw < n_mine;
#if !defined(BLX_ENV_EXPERIMENT_NOFFT) // timing experiment only (wrong results): what the kernel costs without the FFT proper
            fft256_halfwarp<double, BLX_ENV_TW_LOAD != 0>(v, lane16, xchg, tw1, full);
#endif
            __syncwarp(full);
            double2 Bz[8]; // Z[256 - k] for this lane's bins k = lane16 + 16 d
#if BLX_ENV_PARTNER_SHFL
            // Z[256 - k] = Z[(16 - lane16) + 16 (15 - d)] sits in lane 16 - lane16, the register of output 15 - d; lane 0's
            // partners Z[16 (16 - d)] are its own (d = 0: the bin is real, no partner)
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const double2 src = v[fft16_reg_of(15 - d)];
                double2 t;
                t.x = __shfl_sync(full, src.x, (16 - lane16) & 15, 16);
                t.y = __shfl_sync(full, src.y, (16 - lane16) & 15, 16);
                if (d > 0 && lane16 == 0) t = v[fft16_reg_of(16 - d)];
                Bz[d] = t;
            }
#else
#pragma unroll
            for (int r = 0; r < 16; ++r) // only Z[129..255] are read back below
                if (fft16_out_index(r) >= 8) xchg[lane16 + 16 * fft16_out_index(r)] = v[r];
            __syncwarp(full);
#pragma unroll
            for (int d = 0; d < 8; ++d) Bz[d] = xchg[(256 - (lane16 + 16 * d)) & 255];
            __syncwarp(full);
#endif
            double *xr = reinterpret_cast<double *>(xchg); // |X_k|^2, k = 0..256, at pidx(k)
            // bins k = lane16 + 16 d and 256 - k: rows d and 15 - d (lane 0: the last element of rows d - 1 and 15 - d)
            const int ia0 = (lane16 == 0) ? -1 : lane16 - 1; // slot of bin k in row d; lane 0: see below
#if BLX_ENV_TW2_FACTOR
            const double2 wl = tw2[lane16];
#endif
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const int k = lane16 + 16 * d;
                const double2 Zk = v[fft16_reg_of(d)];
                double pa, pb;
                if (k == 0) {
                    const double x0 = Zk.x + Zk.y, xn = Zk.x - Zk.y; // X_0 and X_256 are real
                    pa = 4.0 * (x0 * x0);
                    pb = 4.0 * (xn * xn);
                } else {
#if BLX_ENV_TW2_FACTOR
                    // W512^k = W512^lane16 * W32^d; the second factor is an immediate
                    constexpr double kC32[8] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                                                0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173,
                                                0.19509032201612826785};
                    constexpr double kS32[8] = {0.0, -0.19509032201612826785, -0.38268343236508977173, -0.55557023301960222474,
                                                -0.70710678118654752440, -0.83146961230254523708, -0.92387953251128675613,
                                                -0.98078528040323044913};
                    double2 wk = wl;
                    if (d > 0) {
                        wk.x = wl.x * kC32[d] - wl.y * kS32[d];
                        wk.y = wl.x * kS32[d] + wl.y * kC32[d];
                    }
#else
                    const double2 wk = tw2[k];
#endif
                    const double sr = Zk.x + Bz[d].x, si = Zk.y - Bz[d].y;
                    const double dr = Zk.x - Bz[d].x, di = Zk.y + Bz[d].y;
                    const double tr = dr * wk.x - di * wk.y;
                    const double ti = dr * wk.y + di * wk.x;
                    const double ar = sr + ti, ai = si - tr;
                    const double cr = sr - ti, ci = si + tr;
                    pa = ar * ar + ai * ai;
                    pb = cr * cr + ci * ci;
                }
                // bin k: lanes 1..15 -> row d slot lane16 - 1; lane 0 -> bin 16 d = last element of row d - 1 (bin 0: slot 15 of row 0)
                const int ja = (lane16 != 0) ? kPRow * d + ia0 : (d == 0 ? 15 : kPRow * (d - 1) + pslot15(d - 1));
                // bin 256 - k: lanes 1..15 -> row 15 - d slot 15 - lane16; lane 0 -> bin 16 (16 - d) = last element of row 15 - d
                const int jb = kPRow * (15 - d) + ((lane16 != 0) ? 15 - lane16 : pslot15(15 - d));
                xr[ja] = pa;
                xr[jb] = pb;
            }
            if (lane16 == 0) {
                const double2 Zk = v[fft16_reg_of(8)];
                xr[pidx(128)] = 4.0 * (Zk.x * Zk.x + Zk.y * Zk.y);
            }
            __syncwarp(full);
#if defined(BLX_ENV_EXPERIMENT_NOCHAIN) // timing experiment only (wrong results): what the kernel costs without the accumulation
            const double e = xr[pidx(1 + lane16)] + xr[15];
#else
            const double e = float_chain<BLX_ENV_CHAIN_G>(xr, lane16, active, p.slow_chain != 0);
#endif
            if (active && lane16 == 0) p.energy[sd.env_off + hop] = e;
        }
        __syncwarp(full); // the exchange buffers are free for the next pair's FIR block
    }
}

cudaError_t launch_envelope(const EnvelopeParams &p, int max_hops, int n_songs, cudaStream_t st) {
    static PerDeviceOnce once;
    const cudaError_t e0 = once.run([] {
        cudaError_t e = cudaFuncSetAttribute(envelope_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EG<true>::bytes);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(envelope_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EG<false>::bytes);
        if (e == cudaSuccess && getenv("BLX_DEBUG_OCCUPANCY")) {
            int a = 0, b = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, envelope_kernel<true>, EG<true>::threads, EG<true>::bytes);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, envelope_kernel<false>, EG<false>::threads, EG<false>::bytes);
            fprintf(stderr, "envelope_kernel: %d CTAs/SM of %d threads, %d B (doubled mono); %d of %d threads, %d B (stereo)\n", a,
                    EG<true>::threads, EG<true>::bytes, b, EG<false>::threads, EG<false>::bytes);
        }
        return e;
    });
    if (e0 != cudaSuccess) return e0;
    if (max_hops <= 0) return cudaSuccess;
    if (p.dup) {
        dim3 grid((unsigned)((max_hops + EG<true>::hops_per_cta - 1) / EG<true>::hops_per_cta), (unsigned)n_songs);
        envelope_kernel<true><<<grid, EG<true>::threads, EG<true>::bytes, st>>>(p);
    } else {
        dim3 grid((unsigned)((max_hops + EG<false>::hops_per_cta - 1) / EG<false>::hops_per_cta), (unsigned)n_songs);
        envelope_kernel<false><<<grid, EG<false>::threads, EG<false>::bytes, st>>>(p);
    }
    return cudaGetLastError();
}

} // namespace blx
