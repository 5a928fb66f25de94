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
constexpr int kTilesPerCta = 16;
constexpr int kEnvSamples = (kEnvH + 1) * kHop;  // 4352 stream samples per tile (17 blocks of 256)
constexpr int kMainOut = 16;                     // FIR outputs per thread in the main pass (blocks 0..15)
static_assert(kMainOut * kEnvThreads == kEnvH * kHop, "main pass covers 16 blocks");

constexpr int kOffQ = 0;                                   // short[4352]     TMA staging (raw int16)
constexpr int kOffC = kOffQ + kEnvSamples * 2;             // double[4352]    continuous FIR output, swizzled
constexpr int kOffXchg = kOffC + kEnvSamples * 8;          // double2[16][272] FFT exchange; then |X_k|^2
constexpr int kOffBar = kOffXchg + kEnvH * kXchgElems * 16;
constexpr int kEnvSmem = kOffBar + 16;
static_assert(kEnvSmem <= 115712, "two CTAs per SM");

// The FIR output buffer is addressed in 16-byte cells (two doubles) with an XOR swizzle, so that both
// the FIR stores (a thread owns 8 consecutive cells) and the FFT loads (a half-warp reads 16
// consecutive cells, half-warps 128 cells apart) are bank-conflict free.
__device__ __forceinline__ int cswz(int cell) { return cell ^ ((cell >> 3) & 7); }

// reference include/bandpass_coeffs.h:1-7 — coeffs[0][0..8]; the filter is symmetric.
__device__ __forceinline__ double fir_tap(int k) {
    constexpr double c[9] = {-0.0023470, 0.0044613, -0.0114627, 0.0226382, -0.0405147,
                             0.0580037,  -0.0779167, 0.0882711, 0.9065095};
    return c[k];
}

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

__device__ __forceinline__ double float_chain_scan(const double *xr, int lane16, bool active) {
    const unsigned full = 0xffffffffu;
    const int lane_base = (threadIdx.x & 16);
    double pv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pv[i] = xr[17 * lane16 + 1 + i]; // bins 16 lane16 + 1 + i
    double r = 0.0;
    if (active) { // bins 0..16 one by one (the sum climbs through several binades here)
        float sf = 0.0f;
#pragma unroll
        for (int k = 0; k <= 16; ++k) sf = (float)((double)sf + xr[k]);
        r = (double)sf;
    }
    // Largest remaining bin (by its high word, enough for an exponent test): if it is below 2^25 grid
    // units of the sum after bin 16 it stays so in every later binade: no overflow guard in the rounds.
    int hmax = 0;
    if (lane16 != 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) hmax = max(hmax, __double2hiint(pv[i]));
    }
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) hmax = max(hmax, __shfl_xor_sync(full, hmax, o, 16));
    const bool guard = active && ((hmax >> 20) > ((__double2hiint(r) >> 20) & 0x7ff) + 1);
    int kdone = 16; // bins 0..kdone are in r
    bool done = !active;
    if (__any_sync(full, guard)) {
        while (__any_sync(full, !done)) chain_round<true>(xr, pv, lane16, lane_base, r, kdone, done);
    } else {
        while (__any_sync(full, !done)) chain_round<false>(xr, pv, lane16, lane_base, r, kdone, done);
    }
    return r;
}
} // namespace

// Coefficient that multiplies x[j - m] in the 17-tap FIR, m = 0..16 (symmetric; reference
// include/bandpass_coeffs.h:1-7 and the loop of reference src/tempo_atk_sort.c:124-137).
__device__ __forceinline__ double fir_coef_of_lag(int m) { return fir_tap(m <= 8 ? m : 16 - m); }

// Output t (0..15) of a window whose delay line starts empty (reference src/tempo_atk_sort.c:121):
// xs[m] = raw sample at lag m (already 0 where t - m < 0), summed in the reference's order; the
// affine normalisation only sees the coefficients of the taps inside the window.
__device__ __forceinline__ double head_output(const double (&xs)[16], int t, double A, double Bm) {
    double y = 0;
#pragma unroll
    for (int k = 7; k >= 1; --k) y += fir_tap(k) * (xs[k] + xs[16 - k]);
    y += xs[8] * fir_tap(8);
    y += fir_tap(0) * (xs[0] + 0.0);
    double csum = 0;
#pragma unroll
    for (int m = 0; m < 16; ++m) csum += (m <= t) ? fir_coef_of_lag(m) : 0.0;
    return fma(y, A, -Bm * csum);
}

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
    double2 *ccell = reinterpret_cast<double2 *>(smem + kOffC);
    double2 *xchg_all = reinterpret_cast<double2 *>(smem + kOffXchg);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + kOffBar);

    const int tid = threadIdx.x;
    const int dup = p.dup;
    const short *stream = p.stream + (dup ? sd.q_off : sd.pcm_off);
    auto issue_tile = [&](int t) { // one elected thread
        const int m0 = (tile0 + t) * kEnvH;
        const int h_cnt = min(kEnvH, sd.n_hops - m0);
        const long long base = (long long)m0 * kHop;
        const int n_need = (h_cnt + 1) * kHop; // always inside the song: (m + 2) * 256 <= 512 F <= n
        const unsigned bytes = (unsigned)(dup ? n_need : 2 * n_need);
        const short *src = stream + (dup ? (base >> 1) : base);
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
    // x = (s / 32768 - mean_d) / var_d (reference src/tempo_atk_sort.c:110-113) is affine in the raw
    // sample s, and the FIR is linear: y = A * (sum c_k s_k) - Bm * (sum c_k), A = 1 / (32768 var_d),
    // Bm = mean_d / var_d. The taps run over exact integers; one FMA per output normalises.
    const double A = nm.inv_var_d * (1.0 / 32768);
    const double Bm = nm.mean_d * nm.inv_var_d;
    double csum_all = fir_tap(8);
#pragma unroll
    for (int k = 0; k < 8; ++k) csum_all += 2.0 * fir_tap(k);
    const double Ball = Bm * csum_all;
    unsigned parity = 0;

    for (int t = 0; t < n_tiles; ++t) {
        const int m0 = (tile0 + t) * kEnvH;
        const int h_cnt = min(kEnvH, sd.n_hops - m0);
        mbar_wait(bar, parity);
        parity ^= 1u;

        // ---- continuous FIR, main pass: blocks 0..15, 16 consecutive outputs per thread
        {
            const int j0 = kMainOut * tid; // first output; inputs j0 - 16 .. j0 + 15
            double yo[kMainOut];
            if (dup) {
                // The stream is a mono signal u with every sample doubled (L = R): x[2 i] = x[2 i + 1] = u[i].
                // The 17 taps then fold into two 9-tap filters over u, one for even and one for odd
                // outputs (same products, summed in a different order than the reference's: ~1e-16).
                const int4 *src = reinterpret_cast<const int4 *>(qs + ((j0 - 16) >> 1));
                int wds[8];
                if (tid == 0) { // the delay line is empty before the tile (those outputs are replaced by the heads)
#pragma unroll
                    for (int i = 0; i < 4; ++i) wds[i] = 0;
                    const int4 u1 = reinterpret_cast<const int4 *>(qs)[0];
                    wds[4] = u1.x; wds[5] = u1.y; wds[6] = u1.z; wds[7] = u1.w;
                } else {
                    const int4 u0 = src[0], u1 = src[1];
                    wds[0] = u0.x; wds[1] = u0.y; wds[2] = u0.z; wds[3] = u0.w;
                    wds[4] = u1.x; wds[5] = u1.y; wds[6] = u1.z; wds[7] = u1.w;
                }
                double u[16]; // u[i0 - 8 .. i0 + 7], i0 = j0 / 2
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    u[2 * i] = int_to_double_exact((int)(short)(wds[i] & 0xffff));
                    u[2 * i + 1] = int_to_double_exact(wds[i] >> 16);
                }
#pragma unroll
                for (int o = 0; o < 8; ++o) { // outputs 2 (i0 + o) and 2 (i0 + o) + 1; u index of i0 + o is 8 + o
                    double ye = fir_coef_of_lag(0) * u[8 + o];
                    double yd = (fir_coef_of_lag(0) + fir_coef_of_lag(1)) * u[8 + o];
#pragma unroll
                    for (int m = 1; m <= 8; ++m) {
                        ye = fma(fir_coef_of_lag(2 * m - 1) + fir_coef_of_lag(2 * m), u[8 + o - m], ye);
                        yd = fma(fir_coef_of_lag(2 * m) + (m < 8 ? fir_coef_of_lag(2 * m + 1) : 0.0), u[8 + o - m], yd);
                    }
                    yo[2 * o] = fma(ye, A, -Ball);
                    yo[2 * o + 1] = fma(yd, A, -Ball);
                }
            } else {
                double xv[32];
                const int4 *src = reinterpret_cast<const int4 *>(qs + (j0 - 16));
                int wds[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    int4 u = make_int4(0, 0, 0, 0);
                    if (tid != 0 || i >= 2) u = src[i]; // the delay line is empty before the tile (never used, see heads)
                    wds[4 * i] = u.x; wds[4 * i + 1] = u.y; wds[4 * i + 2] = u.z; wds[4 * i + 3] = u.w;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    xv[2 * i] = int_to_double_exact((int)(short)(wds[i] & 0xffff));
                    xv[2 * i + 1] = int_to_double_exact(wds[i] >> 16);
                }
#pragma unroll
                for (int o = 0; o < kMainOut; ++o) {
                    double y = 0;
#pragma unroll
                    for (int k = 7; k >= 1; --k) y += fir_tap(k) * (xv[o + 16 - k] + xv[o + k]);
                    y += xv[o + 8] * fir_tap(8);
                    y += fir_tap(0) * (xv[o + 16] + xv[o]);
                    yo[o] = fma(y, A, -Ball);
                }
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) ccell[cswz(8 * tid + m)] = make_double2(yo[2 * m], yo[2 * m + 1]);
        }
        // ---- block 16 (second half of the tile's last window): one output per thread
        {
            const int j = kEnvH * kHop + tid;
            double y = 0, xs[17];
#pragma unroll
            for (int i = 0; i < 17; ++i) {
                const int idx = j - 16 + i;
                xs[i] = int_to_double_exact((int)qs[dup ? (idx >> 1) : idx]);
            }
#pragma unroll
            for (int k = 7; k >= 1; --k) y += fir_tap(k) * (xs[16 - k] + xs[k]);
            y += xs[8] * fir_tap(8);
            y += fir_tap(0) * (xs[16] + xs[0]);
            cbuf[2 * cswz(j >> 1) + (j & 1)] = fma(y, A, -Ball);
        }
        __syncthreads();

        // ---- FFT input: 512 FIR outputs of this half-warp's window, the first 16 recomputed with an
        // empty delay line straight from the raw samples (lanes 0..7 own outputs 2 b, 2 b + 1)
        double2 v[16];
#pragma unroll
        for (int a = 0; a < 16; ++a) v[a] = ccell[cswz(128 * w + 16 * a + lane16)];
        {
            // lane t of the half-warp computes head output t; lanes 0..7 then collect outputs 2 b, 2 b + 1
            const int wbase = kHop * w;
            double xs[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) { // raw sample at lag m, 0 before the window starts
                const int i = lane16 - m;
                const int ii = wbase + (i < 0 ? 0 : i);
                const double sdbl = int_to_double_exact((int)qs[dup ? (ii >> 1) : ii]);
                xs[m] = (i < 0) ? 0.0 : sdbl;
            }
            const double h = head_output(xs, lane16, A, Bm);
            const double h0 = __shfl_sync(0xffffffffu, h, (2 * lane16) & 15, 16);
            const double h1 = __shfl_sync(0xffffffffu, h, (2 * lane16 + 1) & 15, 16);
            if (lane16 < 8) v[0] = make_double2(h0, h1);
        }
        __syncthreads(); // FIR buffer and staging buffer are consumed
        if (tid == 0 && t + 1 < n_tiles) issue_tile(t + 1); // prefetch the next tile

        // ---- 16 x (512-point double real FFT + float-accumulated power), one hop per half-warp;
        // both half-warps of a warp run in lockstep (an idle one works on stale data and is ignored)
        if ((w & ~1) < h_cnt) {
            const unsigned full = 0xffffffffu;
            double2 *xchg = xchg_all + w * kXchgElems;
            fft256_halfwarp<double>(v, lane16, xchg, p.tw1, full);
            __syncwarp(full);
#pragma unroll
            for (int r = 0; r < 16; ++r) xchg[lane16 + 16 * fft16_out_index(r)] = v[r];
            __syncwarp(full);
            double2 B[8]; // Z[256 - k] for this lane's bins k = lane16 + 16 d
#pragma unroll
            for (int d = 0; d < 8; ++d) B[d] = xchg[(256 - (lane16 + 16 * d)) & 255];
            __syncwarp(full);
            double *xr = reinterpret_cast<double *>(xchg); // |X_k|^2, k = 0..256, at pbin(k)
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const int k = lane16 + 16 * d;
                const double2 Zk = v[fft16_reg_of(d)];
                if (k == 0) {
                    const double x0 = Zk.x + Zk.y, xn = Zk.x - Zk.y; // X_0 and X_256 are real
                    xr[pbin(0)] = x0 * x0;
                    xr[pbin(256)] = xn * xn;
                } else {
                    const double2 wk = p.tw2[k];
                    const double sr = Zk.x + B[d].x, si = Zk.y - B[d].y;
                    const double dr = Zk.x - B[d].x, di = Zk.y + B[d].y;
                    const double tr = dr * wk.x - di * wk.y;
                    const double ti = dr * wk.y + di * wk.x;
                    const double ar = sr + ti, ai = si - tr;
                    const double cr = sr - ti, ci = si + tr;
                    xr[pbin(k)] = 0.25 * (ar * ar + ai * ai);
                    xr[pbin(256 - k)] = 0.25 * (cr * cr + ci * ci);
                }
            }
            if (lane16 == 0) {
                const double2 Zk = v[fft16_reg_of(8)];
                xr[pbin(128)] = Zk.x * Zk.x + Zk.y * Zk.y;
            }
            __syncwarp(full);
            const double e = float_chain_scan(xr, lane16, w < h_cnt);
            if (w < h_cnt && lane16 == 0) p.energy[sd.env_off + m0 + w] = e;
        }
        // no barrier here: the next tile's FIR only writes buffers every thread is done with
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
