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
// phases (FIR, FFT) of some overlap the latency-bound phase (accumulation chain) of others.
//
// DUP = true: the stream is a mono signal u with every sample doubled (L = R), stored once
// (x[2 i] = x[2 i + 1] = u[i], the decimated stream pass 1 writes for float32 input). The 17 taps
// then fold into two 9-tap filters over u, one for even and one for odd outputs.
//
// The float accumulation s <- (float)((double)s + p_k), k = 0..256, is a chain of 257 dependent
// roundings (~50-75 cycles each on the FP64 + conversion pipes): run naively it idles the SM. It is
// evaluated by the 16 lanes that hold the hop's spectrum:
//   while s stays inside one binade [2^e, 2^(e+1)) its float grid is g = 2^(e-23) and
//   RN_g(s + p) = s + RN_g(p) because s is a multiple of g; so with q = s / g (a 24-bit integer)
//   the chain is q += I_k with I_k = RN_g(p_k) / g, which one double addition p_k + 1.5 * 2^52 * g
//   leaves in the low mantissa word.
// float_chain (below) predicts the binade of every bin from an exact double prefix scan of the
// spectrum, converts every bin once at its predicted grid, takes the few binade crossings in order
// through the reference's own double-add / float-convert step, and verifies every prediction on
// the way; the rare hop that fails is redone by float_chain_rounds / chain_round, the first form of
// the algorithm: one prefix-scan round per binade (every lane converts its 16 bins at the current
// grid, the first bin k* at which q + prefix reaches 2^24 is found with a ballot, bin k* is added
// with the reference step, the scan restarts after k*).
// The result equals the reference's chain except when s + p_k lies within 2^-29 grid units of a
// rounding midpoint (the reference rounds to double, then to float; exact ties follow the parity of
// the magic constant instead of q).
//
// FP64 throughout. FMA contraction and the folded / corrected summation orders differ from the
// reference's (which has no FMA) by ~1e-16 relative in E[m], nine orders of magnitude below the
// 1e-7 onset-count cliff measured in SURVEY.md App. B.
#include "blx_common.cuh"
#include "fft16.cuh"
#include "kernels.h"

namespace blx {

namespace {
constexpr int kEnvThreads = 256;                 // 8 independent warps
constexpr int kEnvWarps = kEnvThreads / 32;
constexpr int kPairsPerWarp = 32;                // a warp owns 64 consecutive hops
constexpr int kHopsPerCta = 2 * kPairsPerWarp * kEnvWarps;
constexpr int kSlotBytes = kHop * 8;             // one block of 256 FIR outputs (128 cells of 16 bytes)

// Shared-memory plan. Per warp: the FFT exchange buffers of its two hops (whose first 4 KB first hold
// the two new FIR blocks of the pair), the carry block, two raw-sample buffers and their mbarriers.
// A raw buffer holds r[i] = stream element blk * (B + 1) - pre + i, B = first block (= hop) of the pair:
// `pre` samples in front, then the two new blocks.
template <bool DUP> struct EG {
    static constexpr int pre = DUP ? 8 : 16;     // raw elements in front of a block that its FIR reads
    static constexpr int blk = DUP ? 128 : 256;  // raw elements per block
    static constexpr int blk_bytes = blk * 2;
    static constexpr int per_thread = blk / 16;  // raw elements whose outputs one lane computes
    static constexpr int taps = DUP ? 8 : 16;    // head-correction terms per lane
    static constexpr int row = taps + 2;         // doubles per head-table row: taps, D, pad
    static constexpr int raw_bytes = ((pre + 2 * blk) * 2 + 63) / 64 * 64;
    static constexpr int w_xchg = 0;                                  // double2[2][272]; FIR blocks B+1, B+2 at +0, +2048
    static constexpr int w_carry = w_xchg + 2 * kXchgElems * 16;      // FIR block B
    static constexpr int w_raw = w_carry + kSlotBytes;                // 2 raw buffers
    static constexpr int w_bar = w_raw + 2 * raw_bytes;               // 2 mbarriers
    static constexpr int w_bytes = (w_bar + 16 + 127) / 128 * 128;
    static constexpr int off_tab = kEnvWarps * w_bytes;               // double[16][row] head table
    static constexpr int bytes = off_tab + 16 * row * 8;
    static_assert(bytes <= 115712, "two CTAs per SM");
};

// reference include/bandpass_coeffs.h:1-7 — coeffs[0][0..8]; the filter is symmetric.
__device__ __forceinline__ double fir_tap(int k) {
    constexpr double c[9] = {-0.0023470, 0.0044613, -0.0114627, 0.0226382, -0.0405147,
                             0.0580037,  -0.0779167, 0.0882711, 0.9065095};
    return c[k];
}
// Coefficient that multiplies x[j - m] in the 17-tap FIR, m = 0..16 (symmetric; reference
// include/bandpass_coeffs.h:1-7 and the loop of reference src/tempo_atk_sort.c:124-137).
__device__ __forceinline__ double fir_coef_of_lag(int m) { return fir_tap(m <= 8 ? m : 16 - m); }
// Folded taps of the doubled mono stream: even output 2 n = sum_m fold_e(m) u[n - m], odd output
// 2 n + 1 = sum_m fold_d(m) u[n - m], m = 0..8.
__device__ __forceinline__ double fold_e(int m) { return m == 0 ? fir_coef_of_lag(0) : fir_coef_of_lag(2 * m - 1) + fir_coef_of_lag(2 * m); }
__device__ __forceinline__ double fold_d(int m) { return fir_coef_of_lag(2 * m) + (m < 8 ? fir_coef_of_lag(2 * m + 1) : 0.0); }
// the same three functions with a run-time argument (table set-up only)
__constant__ double c_fir_half[9] = {-0.0023470, 0.0044613, -0.0114627, 0.0226382, -0.0405147,
                                     0.0580037,  -0.0779167, 0.0882711, 0.9065095};
__device__ __forceinline__ double fir_coef_rt(int m) { return c_fir_half[m <= 8 ? m : 16 - m]; }
__device__ __forceinline__ double fold_e_rt(int m) { return m == 0 ? fir_coef_rt(0) : fir_coef_rt(2 * m - 1) + fir_coef_rt(2 * m); }
__device__ __forceinline__ double fold_d_rt(int m) { return fir_coef_rt(2 * m) + (m < 8 ? fir_coef_rt(2 * m + 1) : 0.0); }

// Where bin k (0..256) of a hop's power spectrum lives in its exchange buffer: lane b reads its bins
// 16 b + 1 .. 16 b + 16 at stride 17 doubles (conflict free); bins 0..16 are at their own index.
__device__ __forceinline__ int pbin(int k) { return k + ((k + 15) >> 4) - 1 + (k == 0); }

// sum_fft of reference src/tempo_atk_sort.c:142-150 for the two hops of a warp, one per half-warp (see
// the header comment). xr[pbin(k)] = |X_k|^2 of this lane's half-warp; `active` is false for a
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
        r = (double)(float)(r + xr[pbin(kdone + 1)]);
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
        const int c = (kk <= kdone) ? increment(xr[pbin(kk)]) : 0;
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
    const int inc1 = (scan && lanes_over && kk > kdone) ? increment(xr[pbin(kk)]) : 0;
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
            r = (double)(float)(sq + xr[pbin(kstar)]); // reference src/tempo_atk_sort.c:147
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
#pragma unroll
    for (int i = 0; i < 16; ++i) pv[i] = xr[17 * lane16 + 1 + i]; // bins 16 lane16 + 1 + i
    double r = r16;
    int kdone = 16; // bins 0..kdone are in r
    bool done = !active;
    while (__any_sync(full, !done)) chain_round<true>(xr, pv, lane16, lane_base, r, kdone, done);
    return r;
}

// Byte offsets of the per-hop scratch behind the 272 doubles of the power spectrum (the hop's FFT exchange
// buffer is 4352 bytes): integer prefix in front of every bin of lanes' bins j = 16 b + i (vector v of lane b at
// 16 (16 v + b)), its total, the predicted exponent at the end, the predicted exponent (top 16 bits of the
// running double sum) in front of every bin, and each lane's mask of binade-crossing bins.
constexpr int kAuxPb = 0, kAuxTot = 1024, kAuxEnd = 1028, kAuxEg = 1040, kAuxMask = 1552;
static_assert(272 * 8 + kAuxMask + 32 <= kXchgElems * 16, "scratch fits behind the spectrum");

// sum_fft of reference src/tempo_atk_sort.c:142-150 for the two hops of a warp, one per half-warp:
// s <- (float)((double)s + p_k), k = 0..256, with p_k = xr[pbin(k)]. `active` is false for a half-warp
// without a hop. Returns (double)sum_fft on every lane of the half-warp.
//
// While s stays inside one binade [2^e, 2^(e+1)) its float grid is g = 2^(e-23) and RN_g(s + p) =
// s + RN_g(p), so with q = s / g the chain is the integer sum q += I_k, I_k = RN_g(p_k) / g, which one
// double addition p_k + 1.5 * 2^52 * g leaves in the low mantissa word. Which binade s is in when bin k
// is added is PREDICTED from the exact prefix sum S_(k-1) of the p_k in double (the float chain stays
// within 257 * 2^-25 relative of it): lane b owns bins 16 b + 1 .. 16 b + 16, a double prefix scan over the
// half-warp gives S, every bin is converted ONCE at its predicted grid, and the bins at which the
// predicted binade changes ("crossings", ~4 per hop) are added one after the other with the reference's
// own double-add / float-convert step, the integer sums of the bins between them coming from an
// integer prefix scan. Bins 0..16 (lane 0), where the sum climbs a binade per bin, run as the plain
// chain meanwhile. Every prediction is VERIFIED on the way (the grid assumed for a run of bins is the
// grid the chain really has; the sum stays below 2^24 grid units up to the next crossing); a hop that
// fails (~6e-5 of random spectra: a prefix sum within rounding noise of a power of two) is redone by the
// binade-by-binade scan above. Host model and test against the sequential chain: tools/chain_model.c.
__device__ __forceinline__ double float_chain(double *xr, int lane16, bool active, bool force_slow) {
    const unsigned full = 0xffffffffu;
    unsigned char *aux = reinterpret_cast<unsigned char *>(xr + 272);
    double pv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pv[i] = xr[17 * lane16 + 1 + i]; // bins 16 lane16 + 1 + i
    // bins 0..16 one by one (independent of everything up to the resolution below)
    double r16;
    {
        float sf = 0.0f;
#pragma unroll
        for (int k = 0; k <= 16; ++k) sf = (float)((double)sf + xr[k]);
        r16 = (double)sf;
    }
    // exact prefix sums in double: local, then across the half-warp
    double c[16];
    c[0] = pv[0] + (lane16 == 0 ? xr[0] : 0.0);
#pragma unroll
    for (int i = 1; i < 16; ++i) c[i] = c[i - 1] + pv[i];
    double inc = c[15];
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
    // every bin at its predicted grid; crossing bins and lane 0's bins (already in r16) count nothing
    const bool mine = active && lane16 != 0;
    int hprev = __double2hiint(excl);
    unsigned acc = 0, mask = 0;
    unsigned L[16], egw[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int hs = __double2hiint(excl + c[i]);
        const int hiM = (hprev & 0x7FF00000) + 0x1D80000; // 1.5 * 2^(e + 29): ulp = float grid of binade e
        const double t = pv[i] + __hiloint2double(hiM, 0);
        const bool x = ((hs ^ hprev) & 0x7FF00000) != 0;
        L[i] = acc;
        acc += (x || !mine) ? 0u : (unsigned)__double2loint(t);
        mask |= x ? (1u << i) : 0u;
        if (i & 1) egw[i >> 1] = __byte_perm(egw[i >> 1], (unsigned)hprev, 0x7610); // high half <- top 16 bits
        else egw[i >> 1] = (unsigned)hprev >> 16;
        hprev = hs;
    }
    if (!mine) mask = 0;
    unsigned incI = acc;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const unsigned up = __shfl_up_sync(full, incI, o, 16);
        if (lane16 >= o) incI += up;
    }
    const unsigned baseI = incI - acc;
#pragma unroll
    for (int v4 = 0; v4 < 4; ++v4)
        *reinterpret_cast<uint4 *>(aux + kAuxPb + 16 * (16 * v4 + lane16)) =
            make_uint4(baseI + L[4 * v4], baseI + L[4 * v4 + 1], baseI + L[4 * v4 + 2], baseI + L[4 * v4 + 3]);
#pragma unroll
    for (int v4 = 0; v4 < 2; ++v4)
        *reinterpret_cast<uint4 *>(aux + kAuxEg + 16 * (16 * v4 + lane16)) =
            make_uint4(egw[4 * v4], egw[4 * v4 + 1], egw[4 * v4 + 2], egw[4 * v4 + 3]);
    *reinterpret_cast<unsigned short *>(aux + kAuxMask + 2 * lane16) = (unsigned short)mask;
    if (lane16 == 15) {
        *reinterpret_cast<unsigned *>(aux + kAuxTot) = incI;
        *reinterpret_cast<int *>(aux + kAuxEnd) = hprev >> 20;
    }
    __syncwarp(full);
    // the crossings in order (no warp-wide operation in here: the two half-warps run their own count)
    int ex = (__double2hiint(r16) >> 20) & 0x7ff;
    unsigned q = (((unsigned)__double2hiint(r16) & 0xFFFFFu) << 3) | ((unsigned)__double2loint(r16) >> 29) | 0x800000u;
    bool ok = !active || (ex >= 1023 - 126 && ex <= 1023 + 126);
    unsigned Pprev = 0;
    const uint4 m0 = *reinterpret_cast<const uint4 *>(aux + kAuxMask), m1 = *reinterpret_cast<const uint4 *>(aux + kAuxMask + 16);
    const unsigned words[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        unsigned m = words[w];
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const int b = 2 * w + (bit >> 4), i = bit & 15;
            const unsigned Pbj = *reinterpret_cast<const unsigned *>(aux + kAuxPb + 16 * (16 * (i >> 2) + b) + 4 * (i & 3));
            const unsigned eg16 = *reinterpret_cast<const unsigned short *>(aux + kAuxEg + 16 * (16 * (i >> 3) + b) + 2 * (i & 7));
            const double pk = xr[17 * b + i + 1];
            const unsigned qb = q + (Pbj - Pprev);
            ok = ok && qb < (1u << 24) && (int)(eg16 >> 4) == ex;
            const double sq = __hiloint2double((ex << 20) | (int)((qb & 0x7FFFFFu) >> 3), (int)((qb & 7u) << 29));
            const double r = (double)(float)(sq + pk); // reference src/tempo_atk_sort.c:147
            ex = (__double2hiint(r) >> 20) & 0x7ff;
            q = (((unsigned)__double2hiint(r) & 0xFFFFFu) << 3) | ((unsigned)__double2loint(r) >> 29) | 0x800000u;
            Pprev = Pbj;
        }
    }
    const unsigned qf = q + (*reinterpret_cast<const unsigned *>(aux + kAuxTot) - Pprev);
    ok = ok && (!active || (qf < (1u << 24) && *reinterpret_cast<const int *>(aux + kAuxEnd) == ex && ex <= 1023 + 126));
    double r = __hiloint2double((ex << 20) | (int)((qf & 0x7FFFFFu) >> 3), (int)((qf & 7u) << 29));
    if (rest_zero) { r = r16; ok = true; }
    if (__any_sync(full, !ok) || force_slow) r = float_chain_rounds(xr, lane16, active, r16);
    return r;
}

} // namespace

template <bool DUP> __global__ void __launch_bounds__(kEnvThreads, 2) envelope_kernel(EnvelopeParams p) {
    using G = EG<DUP>;
    extern __shared__ __align__(128) unsigned char smem[];
    const SongDesc sd = p.songs[blockIdx.y];
    if (blockIdx.x * kHopsPerCta >= sd.n_hops) return;
    const SongNorm nm = p.norm[blockIdx.y];
    if (nm.status != 0) return;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int hw = lane >> 4, lane16 = lane & 15; // half-warp = hop of the pair
    double *tab = reinterpret_cast<double *>(smem + G::off_tab);
    // Head table: a window's output t (its first 16) lacks the terms of the `pre` raw samples X[0..pre)
    // in front of the window; row `lane16` holds the coefficients of those terms for the output this
    // lane corrects, then the sum of the dropped coefficients (for the mean term).
    //   DUP: lane = n + 8 part, output 2 n + part: sum_{j >= n} fold(8 + n - j) X[j]
    //   else: lane = output j:                      sum_{i >= j} c(j + 16 - i) X[i]
    if (tid < 16) {
        double *trow = tab + tid * G::row;
        double dsum = 0.0;
        if (DUP) {
            const int n = tid & 7, part = tid >> 3;
            for (int j = 0; j < 8; ++j) {
                const int m = 8 + n - j;
                const double c = (j >= n) ? (part ? fold_d_rt(m) : fold_e_rt(m)) : 0.0;
                trow[j] = c;
                dsum += c;
            }
        } else {
            for (int i = 0; i < 16; ++i) {
                const double c = (i >= tid) ? fir_coef_rt(tid + 16 - i) : 0.0;
                trow[i] = c;
                dsum += c;
            }
        }
        trow[G::taps] = dsum;
        trow[G::taps + 1] = 0.0;
    }

    unsigned char *wsm = smem + warp * G::w_bytes;
    double2 *xchg = reinterpret_cast<double2 *>(wsm + G::w_xchg) + hw * kXchgElems;
    unsigned char *newblk = wsm + G::w_xchg; // FIR blocks B + 1, B + 2 of the pair (consumed before the FFT exchange)
    unsigned char *carry = wsm + G::w_carry; // FIR block B
    unsigned char *raw = wsm + G::w_raw;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wsm + G::w_bar);

    const int B0 = blockIdx.x * kHopsPerCta + warp * (2 * kPairsPerWarp); // first hop (= first block) of this warp
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

    // x = (s / 32768 - mean_d) / var_d (reference src/tempo_atk_sort.c:110-113) is affine in the raw
    // sample s, and the FIR is linear: y = A * (sum c_k s_k) - Bm * (sum c_k), A = 1 / (32768 var_d),
    // Bm = mean_d / var_d. The taps run over exact integers; one FMA per output normalises.
    // Both carry an extra factor 1/2 (exact): it is the 1/2 of the real-FFT even/odd split, so the split
    // below produces X_k without its 0.25 |.|^2 scaling step; the three purely real bins are scaled back.
    const double A = nm.inv_var_d * (1.0 / 32768) * 0.5;
    const double Bm = nm.mean_d * nm.inv_var_d * 0.5;
    double csum_all = fir_tap(8);
#pragma unroll
    for (int k = 0; k < 8; ++k) csum_all += 2.0 * fir_tap(k);
    const double Ball = Bm * csum_all;
    const unsigned full = 0xffffffffu;
    const int key = 16 * (lane16 & 7);
    // byte offset of ring cell 16 a + lane16 inside a block: 256 a + 16 (lane16 ^ ((2 a + (lane16 >> 3)) & 7))
    int coff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) coff[i] = 16 * (lane16 ^ ((2 * i + (lane16 >> 3)) & 7));

    for (int q = -1; q < n_pairs; ++q) {
        const unsigned char *rcur = raw + (q & 1) * G::raw_bytes;
        mbar_wait(bar + (q & 1), (unsigned)((q + 1) >> 1) & 1u);

        // ---- continuous FIR: half-warp hw computes block B + 1 + hw, every lane 16 consecutive outputs
        // (8 cells); in the prologue only the upper half-warp's block (B0) is kept
        {
            const unsigned char *sp = rcur + lane * (G::per_thread * 2); // r[per_thread * lane ...]
            double yo[16];
            if (DUP) {
                // inputs u[k] = r[8 lane + k], k = 0..15; output pair o uses u[8 + o - m], m = 0..8
                const int4 u0 = reinterpret_cast<const int4 *>(sp)[0], u1 = reinterpret_cast<const int4 *>(sp)[1];
                const int wds[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
                double u[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    u[2 * i] = (double)(short)(wds[i] & 0xffff);
                    u[2 * i + 1] = (double)(short)(wds[i] >> 16);
                }
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    double ye = fold_e(0) * u[8 + o];
                    double yd = fold_d(0) * u[8 + o];
#pragma unroll
                    for (int m = 1; m <= 8; ++m) {
                        ye = fma(fold_e(m), u[8 + o - m], ye);
                        yd = fma(fold_d(m), u[8 + o - m], yd);
                    }
                    yo[2 * o] = fma(ye, A, -Ball);
                    yo[2 * o + 1] = fma(yd, A, -Ball);
                }
            } else {
                // inputs xv[k] = r[16 lane + k], k = 0..31; output o uses xv[o + 16 - m], m = 0..16, summed in
                // the reference's order (reference src/tempo_atk_sort.c:124-137)
                double xv[32];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int4 q4 = reinterpret_cast<const int4 *>(sp)[i];
                    const int wds[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        xv[8 * i + 2 * j] = (double)(short)(wds[j] & 0xffff);
                        xv[8 * i + 2 * j + 1] = (double)(short)(wds[j] >> 16);
                    }
                }
#pragma unroll
                for (int o = 0; o < 16; ++o) {
                    double y = 0;
#pragma unroll
                    for (int k = 7; k >= 1; --k) y += fir_tap(k) * (xv[o + 16 - k] + xv[o + k]);
                    y += xv[o + 8] * fir_tap(8);
                    y += fir_tap(0) * (xv[o + 16] + xv[o]);
                    yo[o] = fma(y, A, -Ball);
                }
            }
            unsigned char *slot = ((q < 0) ? carry : newblk + hw * kSlotBytes) + 128 * lane16;
            if (q >= 0 || hw == 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m)
                    *reinterpret_cast<double2 *>(slot + ((16 * m) ^ key)) = make_double2(yo[2 * m], yo[2 * m + 1]);
            }
        }
        __syncwarp(full);
        if (q < 0) continue;

        // ---- FFT input: the 512 FIR outputs of this half-warp's window: blocks B, B + 1 (hop 0) or
        // B + 1, B + 2 (hop 1) ...
        double2 v[16];
        {
            const unsigned char *ba = hw ? newblk : carry, *bb = newblk + hw * kSlotBytes;
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                v[a] = *reinterpret_cast<const double2 *>(ba + 256 * a + coff[a & 3]);
                v[a + 8] = *reinterpret_cast<const double2 *>(bb + 256 * a + coff[a & 3]);
            }
        }
        // ... whose first 16 see an empty delay line: take out the terms of the raw samples in front of
        // the window (head table above) and put back the mean term of the dropped coefficients. Hop 1's
        // window starts at r[pre]; hop 0's one block earlier, at r'[blk + pre] of the previous pass.
        {
            const unsigned char *hp = hw ? rcur : raw + ((q & 1) ^ 1) * G::raw_bytes + G::blk_bytes;
            const double *trow = tab + lane16 * G::row;
            double X[G::taps];
#pragma unroll
            for (int i = 0; i < G::taps / 8; ++i) {
                const int4 q4 = reinterpret_cast<const int4 *>(hp)[i];
                const int wds[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    X[8 * i + 2 * j] = (double)(short)(wds[j] & 0xffff);
                    X[8 * i + 2 * j + 1] = (double)(short)(wds[j] >> 16);
                }
            }
            double corr = 0.0;
#pragma unroll
            for (int i = 0; i < G::taps; i += 2) {
                const double2 c2 = *reinterpret_cast<const double2 *>(trow + i);
                corr = fma(c2.x, X[i], corr);
                corr = fma(c2.y, X[i + 1], corr);
            }
            const double h = fma(Bm, trow[G::taps], -(A * corr));
            const int s0 = DUP ? lane16 : 2 * lane16, s1 = DUP ? lane16 + 8 : 2 * lane16 + 1;
            const double h0 = __shfl_sync(full, h, s0 & 15, 16);
            const double h1 = __shfl_sync(full, h, s1 & 15, 16);
            if (lane16 < 8) { v[0].x += h0; v[0].y += h1; }
        }
        __syncwarp(full); // FIR blocks and raw samples are consumed
        if (hw == 1) { // block B + 2 (this half-warp's second block) becomes the next pair's block B
#pragma unroll
            for (int a = 0; a < 8; ++a) *reinterpret_cast<double2 *>(carry + 256 * a + coff[a & 3]) = v[a + 8];
        }
        if (lane == 0 && q + 1 < n_pairs) issue_pass(q + 1);

        // ---- 2 x (512-point double real FFT + float-accumulated power), one hop per half-warp in
        // lockstep (a half-warp without a hop works on stale data and is ignored)
        {
            const int hop = B0 + 2 * q + hw;
            const bool active = 2 * q + hw < n_mine;
            fft256_halfwarp<double>(v, lane16, xchg, p.tw1, full);
            __syncwarp(full);
#pragma unroll
            for (int r = 0; r < 16; ++r) // only Z[129..255] are read back below
                if (fft16_out_index(r) >= 8) xchg[lane16 + 16 * fft16_out_index(r)] = v[r];
            __syncwarp(full);
            double2 Bz[8]; // Z[256 - k] for this lane's bins k = lane16 + 16 d
#pragma unroll
            for (int d = 0; d < 8; ++d) Bz[d] = xchg[(256 - (lane16 + 16 * d)) & 255];
            __syncwarp(full);
            double *xr = reinterpret_cast<double *>(xchg); // |X_k|^2, k = 0..256, at pbin(k)
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const int k = lane16 + 16 * d;
                const double2 Zk = v[fft16_reg_of(d)];
                if (k == 0) {
                    const double x0 = Zk.x + Zk.y, xn = Zk.x - Zk.y; // X_0 and X_256 are real
                    xr[pbin(0)] = 4.0 * (x0 * x0);
                    xr[pbin(256)] = 4.0 * (xn * xn);
                } else {
                    const double2 wk = p.tw2[k];
                    const double sr = Zk.x + Bz[d].x, si = Zk.y - Bz[d].y;
                    const double dr = Zk.x - Bz[d].x, di = Zk.y + Bz[d].y;
                    const double tr = dr * wk.x - di * wk.y;
                    const double ti = dr * wk.y + di * wk.x;
                    const double ar = sr + ti, ai = si - tr;
                    const double cr = sr - ti, ci = si + tr;
                    xr[pbin(k)] = ar * ar + ai * ai;
                    xr[pbin(256 - k)] = cr * cr + ci * ci;
                }
            }
            if (lane16 == 0) {
                const double2 Zk = v[fft16_reg_of(8)];
                xr[pbin(128)] = 4.0 * (Zk.x * Zk.x + Zk.y * Zk.y);
            }
            __syncwarp(full);
            const double e = float_chain(xr, lane16, active, p.slow_chain != 0);
            if (active && lane16 == 0) p.energy[sd.env_off + hop] = e;
        }
        __syncwarp(full); // the exchange buffers are free for the next pair's FIR blocks
    }
}

cudaError_t launch_envelope(const EnvelopeParams &p, int max_hops, int n_songs, cudaStream_t st) {
    static PerDeviceOnce once;
    if (once.first_time()) {
        cudaError_t e = cudaFuncSetAttribute(envelope_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EG<true>::bytes);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(envelope_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EG<false>::bytes);
        if (e != cudaSuccess) return e;
    }
    if (max_hops <= 0) return cudaSuccess;
    dim3 grid((unsigned)((max_hops + kHopsPerCta - 1) / kHopsPerCta), (unsigned)n_songs);
    if (p.dup) envelope_kernel<true><<<grid, kEnvThreads, EG<true>::bytes, st>>>(p);
    else envelope_kernel<false><<<grid, kEnvThreads, EG<false>::bytes, st>>>(p);
    return cudaGetLastError();
}

} // namespace blx
