// pass1.cu — the streaming first pass over a song's PCM (kernel id BLX_K_PASS1).
//
// One launch covers a batch of songs. A CTA (128 threads) owns one (song, part) pair and walks
// tiles of 8 frames (4096 per-channel samples) of that song:
//
//   1. TMA. Float32 input (the benchmark path): the packed PCM buffer is described by two 2-D tensor maps
//      over rows of 64 floats (256 bytes, one row per thread) - the first 128 bytes of every row and the
//      second 128 bytes - and ONE elected thread issues two tensor copies per tile (cp.async.bulk.tensor,
//      SASS UTMALDG), 129 rows each (the tile's 128 rows plus the half row of FIR history / look-ahead on
//      either side), with the 128-byte swizzle, so that the per-thread 128-bit reads of the next step are
//      bank-conflict free without padding. int16 input: every thread issues one 1-D bulk copy (UBLKCP) of
//      its row into a padded shared row. Completion is tracked by one mbarrier per CTA; the copy for the
//      next tile is issued as soon as the rows are consumed, so it overlaps the FFTs.
//   2. Per row: front-end (F32 input: 23-tap half-band 2:1 decimation + int16 quantisation, bit-exact
//      with blx_frontend.h) or stereo down-mix (S16 input: (L + R) / 2, C truncation, reference
//      src/frequency_sort.c:71-74), times the Hann window (reference src/frequency_sort.c:40-42,74),
//      written to the FFT staging buffer. In FULL mode the same registers feed
//        - the exact integer histogram of sample values -1904..+1902 (the only bins that can reach the
//          integral of reference src/amplitude_sort.c:69-71 after 301 smoothing passes),
//        - sum / sum of squares / first and last non-zero index (reference src/helpers.c:30-49,
//          src/amplitude_sort.c:26-31),
//        - for F32 input, the decimated int16 stream consumed by the envelope pass.
//   3. Eight 512-point real FFTs at once, 16 threads each (fft16.cuh), and per-bin |X_d|^2 accumulated
//      in registers over all tiles of the CTA (reference src/frequency_sort.c:83-93).
//
// At the end the eight per-group accumulators are reduced in a fixed order into one partial spectrum
// per CTA; the epilogue kernel (epilogue.cu) sums the partials of a song in order, so results are
// deterministic run to run.
#include <type_traits>

#include "blx_common.cuh"
#include "fft16.cuh"
#include "kernels.h"

namespace blx {

namespace {

constexpr int kP1Threads = 128;
constexpr int kP1FramesPerTile = 8;
constexpr int kP1TileM = kP1FramesPerTile * kWin; // 4096 per-channel samples per tile
constexpr int kFinFrame = 16 * 36;                // floats per staged frame: 16 chunks of 32 (+4 pad)

template <int KIND> struct RowGeom;
template <> struct RowGeom<kInF32> { static constexpr int elems = 64, ebytes = 4, halo = 1; };
template <> struct RowGeom<kInS16Stereo> { static constexpr int elems = 64, ebytes = 2, halo = 0; };
template <> struct RowGeom<kInS16Mono> { static constexpr int elems = 32, ebytes = 2, halo = 0; };

template <int KIND> struct P1Smem {
    using G = RowGeom<KIND>;
    static constexpr int row_bytes = G::elems * G::ebytes;
    static constexpr int row_stride = row_bytes + 16;
    static constexpr int n_slots = kP1Threads + 2 * G::halo;
    // F32: region A = first halves of rows 0..128 (129 x 128 bytes), region B = second halves of rows -1..127,
    // both written by tensor copies with the 128-byte swizzle; B starts on a 1024-byte boundary.
    static constexpr int f32_half = 128, f32_rows = kP1Threads + 1;
    static constexpr int off_a = 0, off_b = 17 * 1024;
    static constexpr int raw_bytes = (KIND == kInF32) ? off_b + f32_rows * f32_half : n_slots * row_stride;
    static constexpr int off_raw = 0;
    static constexpr int off_fin = (raw_bytes + 127) / 128 * 128;
    static constexpr int fin_bytes = kP1FramesPerTile * kFinFrame * 4; // 18432; also 8 KB reduction scratch
    static constexpr int off_hann = off_fin + fin_bytes;
    static constexpr int off_tw1 = off_hann + kFinFrame * 4;
    static constexpr int off_bar = off_tw1 + 256 * 8;
    static constexpr int off_hist = off_bar + 16;
    static constexpr int bytes_lite = off_hist;
    static constexpr int bytes_full = off_hist + (kHistStride + 32) * 4; // + one spare counter per lane (hist_add)
};

// 2-D tensor copy global -> shared (SASS: UTMALDG), completion on an mbarrier. c0 = element, c1 = row.
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
// Address of 16-byte chunk c of the 128-byte row at `row` inside a swizzled region (row is 128-byte aligned):
// the copy engine XORs the chunk index with bits 7..9 of the shared-memory address.
__device__ __forceinline__ const float4 *swz_chunk(const unsigned char *row, unsigned key16, int c) {
    return reinterpret_cast<const float4 *>(row + (((unsigned)c << 4) ^ key16));
}
__device__ __forceinline__ unsigned swz_key16(const void *row) { return ((smem_u32(row) >> 7) & 7u) << 4; }

// Histogram of sample values -1904..+1902 (the only bins that can reach the integral of reference
// src/amplitude_sort.c:69-71): ONE UNCONDITIONAL shared-memory reduction per sample on a 32-bit shared-window address;
// a sample outside the range adds to a spare counter of its lane behind the bins (never flushed). Measured on B200
// per 1 024 benchmark songs (a third of the samples out of range): C++ `if (bin < n) atomicAdd(&hist[bin], c)` 11.1 ms
// (divergent branch + generic-to-shared address conversion per sample), predicated `red.shared` on a shared-window
// address 10.3 ms (ptxas still branches around it), unconditional with ONE shared spare counter 11.8 ms (same-address
// reductions serialise), unconditional with a spare counter per lane 9.3 ms.
__device__ __forceinline__ void hist_add(unsigned hist_s32, int v, unsigned c) {
    const unsigned bin = (unsigned)(v + 32768 - kHistLo);
    const unsigned slot = (bin < (unsigned)kHistBins) ? bin : (unsigned)kHistStride + (threadIdx.x & 31u);
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hist_s32 + 4u * slot), "r"(c) : "memory");
}

struct ThreadStats {
    long long sum;            // sum of samples            (bl_mean, reference src/helpers.c:30-37)
    unsigned long long sumsq; // sum of squared samples    (bl_variance, reference src/helpers.c:39-49)
    int first_nz, last_nz;    // reference src/amplitude_sort.c:26-31
};
} // namespace

#ifndef BLX_P1_TW_REG
// 1: the 15 inter-pass FFT twiddles of a thread stay in registers for the life of the CTA (30 registers) instead of being
// read from shared memory for every frame; the register allocation is sized for the CTAs per SM that shared memory allows
#define BLX_P1_TW_REG 1
#endif
#if BLX_P1_TW_REG
#define BLX_P1_BOUNDS __launch_bounds__(kP1Threads, (KIND == kInF32 ? 3 : 4) + (FULL ? 0 : 1)) // = what P1Smem<KIND> lets an SM hold
#else
#define BLX_P1_BOUNDS __launch_bounds__(kP1Threads)
#endif

template <int KIND, bool FULL>
__global__ void BLX_P1_BOUNDS pass1_kernel(const __grid_constant__ Pass1Params p) {
    using SM = P1Smem<KIND>;
    using G = RowGeom<KIND>;
    extern __shared__ __align__(128) unsigned char smem[];
    const SongDesc sd = p.songs[blockIdx.y];
    const int part = blockIdx.x;
    if (part >= sd.n_parts) return;

    unsigned char *raw = smem + SM::off_raw;
    float *fin = reinterpret_cast<float *>(smem + SM::off_fin);
    float *hannp = reinterpret_cast<float *>(smem + SM::off_hann);
    float2 *tw1 = reinterpret_cast<float2 *>(smem + SM::off_tw1);
    const float2 *tw2 = p.tw2; // 8 entries per thread and tile: read through L1 (keeps four CTAs per SM in shared memory)
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + SM::off_bar);
    unsigned *hist = reinterpret_cast<unsigned *>(smem + SM::off_hist);
    const unsigned hist_s32 = smem_u32(hist);

    const int tid = threadIdx.x;
    const int lane16 = tid & 15;
    const int grp = tid >> 4; // frame slot inside the tile == FFT group

    // ---- one-time setup: tables into shared memory, histogram cleared, barrier armed
    for (int i = tid; i < kWin; i += kP1Threads) hannp[(i >> 5) * 36 + (i & 31)] = p.hann[i];
    for (int i = tid; i < 256; i += kP1Threads) tw1[i] = p.tw1[i];
    if (FULL)
        for (int i = tid; i < kHistStride; i += kP1Threads) hist[i] = 0u;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
#if BLX_P1_TW_REG
    float2 twr[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) twr[c] = p.tw1[c * 16 + lane16];
#endif
    __syncthreads();

    const unsigned char *src = reinterpret_cast<const unsigned char *>(p.pcm) + (size_t)sd.pcm_off * G::ebytes;
    const long long n_elems = sd.n_elems;

    // Issues the bulk copies of one tile. Row `row` (may be -1 or 128 for the halo) starts at input
    // element tile * 128 * elems + row * elems; it is copied in full iff it starts inside the song
    // (the buffer is readable up to the next multiple of BLX_ALIGN_ELEMS).
    auto issue_tile = [&](int tile) {
        const long long e0 = (long long)tile * kP1Threads * G::elems;
        fence_proxy_async();
        if (KIND == kInF32) {
            // rows of 64 floats; the tile's rows 0..128 (first halves) and -1..127 (second halves). Rows outside
            // the buffer arrive as zeros, rows of a neighbouring song are cleared by the fix-up below.
            if (tid == 0) {
                const int row0 = (int)(sd.pcm_off >> 6) + tile * kP1Threads;
                mbar_arrive_expect_tx(bar, 2u * SM::f32_rows * SM::f32_half);
                tma_load_2d(raw + SM::off_a, &p.map_a, 0, row0, bar);
                tma_load_2d(raw + SM::off_b, &p.map_b, 0, row0 - 1, bar);
            }
            return;
        }
        if (tid == 0) {
            long long first = e0 - (long long)G::halo * G::elems;
            if (first < 0) first = 0;
            long long last = e0 + (long long)(kP1Threads + G::halo) * G::elems; // exclusive
            const long long lim = (n_elems + G::elems - 1) / G::elems * G::elems;
            if (last > lim) last = lim;
            const long long rows = (last > first) ? (last - first) / G::elems : 0;
            mbar_arrive_expect_tx(bar, (unsigned)(rows * SM::row_bytes));
        }
        {
            const long long es = e0 + (long long)tid * G::elems;
            if (es < n_elems)
                tma_load_1d(raw + (size_t)(tid + G::halo) * SM::row_stride, src + es * G::ebytes, SM::row_bytes, bar);
        }
    };

    float accA[8], accB[8], acc128 = 0.0f;
#pragma unroll
    for (int d = 0; d < 8; ++d) accA[d] = accB[d] = 0.0f;

    ThreadStats ts;
    ts.sum = 0; ts.sumsq = 0ull; ts.first_nz = 0x7fffffff; ts.last_nz = -1;

    unsigned parity = 0;
    int tile = part;
    if (tile < sd.n_tiles) issue_tile(tile);

    for (; tile < sd.n_tiles; tile += sd.n_parts) {
        mbar_wait(bar, parity);
        parity ^= 1u;

        const long long e0 = (long long)tile * kP1Threads * G::elems;
        // ---- boundary fix-up (block-uniform): rows that were not copied, and the part of the last
        // row beyond the song, must read as zero (x[i] = 0 outside [0, n_in), blx_frontend.h).
        const bool edge = (e0 - (long long)G::halo * G::elems < 0) ||
                          (e0 + (long long)(kP1Threads + G::halo) * G::elems > n_elems);
        if (edge) {
            for (int slot = tid; slot < SM::n_slots; slot += kP1Threads) {
                const long long es = e0 + (long long)(slot - G::halo) * G::elems;
                if (KIND == kInF32) {
                    // row s = slot - 1: elements 0..31 in region A row s (s >= 0), 32..63 in region B row s + 1 (s <= 127)
                    const int s = slot - 1;
                    const int keep = (es < 0 || es >= n_elems) ? 0 : (int)min((long long)G::elems, n_elems - es);
                    for (int el = keep; el < G::elems; ++el) {
                        const int half = el >> 5;
                        if ((half == 0 && s < 0) || (half == 1 && s > kP1Threads - 1)) continue;
                        unsigned char *row = raw + (half ? SM::off_b + (s + 1) * SM::f32_half : SM::off_a + s * SM::f32_half);
                        *reinterpret_cast<float *>(row + ((((unsigned)(el & 31) >> 2) << 4) ^ swz_key16(row)) + 4 * (el & 3)) = 0.0f;
                    }
                    continue;
                }
                unsigned char *row = raw + (size_t)slot * SM::row_stride;
                if (es < 0 || es >= n_elems) {
                    for (int i = 0; i < SM::row_bytes / 16; ++i) reinterpret_cast<int4 *>(row)[i] = make_int4(0, 0, 0, 0);
                } else if (es + G::elems > n_elems) {
                    const int keep = (int)(n_elems - es);
                    for (int i = keep; i < G::elems; ++i) reinterpret_cast<short *>(row)[i] = 0;
                }
            }
            __syncthreads();
        }

        // ---- per-row front-end / down-mix + window -> fin, plus FULL-mode side products
        const long long m0 = (long long)tile * kP1TileM + tid * 32; // first per-channel sample of this row
        float *fout = fin + grp * kFinFrame + lane16 * 36;
        const float *hrow = hannp + lane16 * 36;

        if (KIND == kInF32) {
            const float H[BLX_FE_NPAIRS] = BLX_FE_TAPS;
            // chunk c (16 bytes) of this thread's row: 0..7 in region A row tid, 8..15 in region B row tid + 1;
            // the previous row's chunks 13..15 are region B row tid, the next row's chunks 0..2 region A row tid + 1
            const unsigned char *ra = raw + SM::off_a + tid * SM::f32_half, *rb = raw + SM::off_b + tid * SM::f32_half;
            const unsigned ka = swz_key16(ra), ka1 = swz_key16(ra + SM::f32_half);
            const unsigned kb = swz_key16(rb), kb1 = swz_key16(rb + SM::f32_half);
            auto own = [&](int c) -> float4 { return c < 8 ? *swz_chunk(ra, ka, c) : *swz_chunk(rb + SM::f32_half, kb1, c - 8); };
            const long long n_out = n_elems >> 1;
            float w[32];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float4 a = *swz_chunk(rb, kb, 5 + i), b = own(i);
                w[4 * i] = a.x; w[4 * i + 1] = a.y; w[4 * i + 2] = a.z; w[4 * i + 3] = a.w;
                w[12 + 4 * i] = b.x; w[12 + 4 * i + 1] = b.y; w[12 + 4 * i + 2] = b.z; w[12 + 4 * i + 3] = b.w;
            }
            const int valid = (int)max(0ll, min(32ll, n_out - m0)); // samples of this row inside the song
            // The row in two compiled forms: CHECK = false for a row entirely inside the song (no per-sample
            // test), CHECK = true for the one row a song ends in (and the rows behind it).
            auto do_row = [&](auto check_tag) {
                constexpr bool CHECK = decltype(check_tag)::value;
                int qw[16]; // the row's 32 int16 values, packed in pairs (what goes to the decimated stream)
                int row_sum = 0;
                unsigned sq_lo = 0, sq_hi = 0; // 64-bit sum of squares
                int first_i = 64, last_i = -1;  // first / last non-zero sample of the row (CHECK rows, or zero ends)
#pragma unroll
                for (int s = 0; s < 8; ++s) {
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const int q4 = 2 * s + 3 + h2;
                        const float4 a = (q4 < 16) ? own(q4) : *swz_chunk(ra + SM::f32_half, ka1, q4 - 16);
                        w[24 + 4 * h2] = a.x; w[25 + 4 * h2] = a.y; w[26 + 4 * h2] = a.z; w[27 + 4 * h2] = a.w;
                    }
                    float o[4];
                    int qi[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = 12 + 2 * u;
                        float acc = __fmul_rn(BLX_FE_CENTER, w[c]);
#pragma unroll
                        for (int k = 0; k < BLX_FE_NPAIRS; ++k)
                            acc = __fmaf_rn(H[k], __fadd_rn(w[c - (2 * k + 1)], w[c + (2 * k + 1)]), acc);
                        float q = rintf(acc);
                        q = fminf(fmaxf(q, -32768.0f), 32767.0f);
                        o[u] = q;
                        if (FULL) qi[u] = (int)q;
                    }
                    if (FULL) {
                        qw[2 * s] = (qi[0] & 0xffff) | (qi[1] << 16);
                        qw[2 * s + 1] = (qi[2] & 0xffff) | (qi[3] << 16);
                        if (!CHECK) {
                            // mono sample -> L = R: every value counts twice (doubled below / in the histogram)
                            row_sum += (qi[0] + qi[1]) + (qi[2] + qi[3]);
                            const unsigned a = (unsigned)(qi[0] * qi[0]) + (unsigned)(qi[1] * qi[1]); // <= 2^31
                            const unsigned b = (unsigned)(qi[2] * qi[2]) + (unsigned)(qi[3] * qi[3]);
                            const unsigned t0 = sq_lo + a;
                            sq_hi += (t0 < a);
                            sq_lo = t0 + b;
                            sq_hi += (sq_lo < b);
#pragma unroll
                            for (int u = 0; u < 4; ++u) hist_add(hist_s32, qi[u], 2u);
                        } else {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (4 * s + u < valid) {
                                    row_sum += qi[u];
                                    const unsigned a = (unsigned)(qi[u] * qi[u]);
                                    sq_lo += a;
                                    sq_hi += (sq_lo < a);
                                    hist_add(hist_s32, qi[u], 2u);
                                    if (qi[u] != 0) { first_i = min(first_i, 4 * s + u); last_i = 4 * s + u; }
                                }
                            }
                        }
                    }
                    const float4 hv = *reinterpret_cast<const float4 *>(hrow + 4 * s);
                    float4 r4;
                    r4.x = o[0] * hv.x; r4.y = o[1] * hv.y; r4.z = o[2] * hv.z; r4.w = o[3] * hv.w;
                    *reinterpret_cast<float4 *>(fout + 4 * s) = r4;
#pragma unroll
                    for (int i = 0; i < 24; ++i) w[i] = w[i + 8];
                }
                if (FULL && (!CHECK || valid > 0)) {
                    ts.sum += 2 * (long long)row_sum;
                    ts.sumsq += 2ull * (((unsigned long long)sq_hi << 32) | sq_lo);
                    // first / last non-zero sample (interleaved index 2 t, 2 t + 1): the row's end samples decide
                    // unless one of them is zero (then look through the row)
                    if (!CHECK) {
                        if ((qw[0] & 0xffff) != 0 && (qw[15] >> 16) != 0) {
                            first_i = 0;
                            last_i = 31;
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const int v = (i & 1) ? (qw[i >> 1] >> 16) : (qw[i >> 1] & 0xffff);
                                if (v != 0) { first_i = min(first_i, i); last_i = i; }
                            }
                        }
                    }
                    if (last_i >= 0) {
                        ts.first_nz = min(ts.first_nz, (int)(2 * (m0 + first_i)));
                        ts.last_nz = max(ts.last_nz, (int)(2 * (m0 + last_i) + 1));
                    }
                    // decimated stream for the envelope pass (mono: L == R): two 256-bit stores (STG.256: a full
                    // 32-byte sector per lane); q rows are padded to 32 samples by the engine
                    short *dst = p.qout + sd.q_off + m0;
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 16 * i), "r"(qw[8 * i]),
                                     "r"(qw[8 * i + 1]), "r"(qw[8 * i + 2]), "r"(qw[8 * i + 3]), "r"(qw[8 * i + 4]), "r"(qw[8 * i + 5]),
                                     "r"(qw[8 * i + 6]), "r"(qw[8 * i + 7])
                                     : "memory");
                }
            };
            if (!FULL || valid == 32) do_row(std::false_type{});
            else do_row(std::true_type{});
        } else {
            const int4 *rowp = reinterpret_cast<const int4 *>(raw + (size_t)tid * SM::row_stride);
            constexpr int per_vec = (KIND == kInS16Stereo) ? 4 : 8; // per-channel samples per 16-byte load
            const long long i0 = (long long)tile * kP1Threads * G::elems + (long long)tid * G::elems;
            const int valid_e = (int)max(0ll, min((long long)G::elems, n_elems - i0)); // elements inside the song
            // as for float input: CHECK = false for a row entirely inside the song, true for the row a song ends in
            auto do_row = [&](auto check_tag) {
                constexpr bool CHECK = decltype(check_tag)::value;
                int row_sum = 0, first_v = 0, last_v = 0;
                unsigned sq_lo = 0, sq_hi = 0; // 64-bit sum of squares
#pragma unroll
                for (int vix = 0; vix < SM::row_bytes / 16; ++vix) {
                    const int4 raw4 = rowp[vix];
                    const int wds[4] = {raw4.x, raw4.y, raw4.z, raw4.w};
                    float o[per_vec];
#pragma unroll
                    for (int wi = 0; wi < 4; ++wi) {
                        const int lo = (int)(short)(wds[wi] & 0xffff);
                        const int hi = wds[wi] >> 16;
                        if (KIND == kInS16Stereo) o[wi] = (float)((lo + hi) / 2); // C truncation toward zero
                        else { o[2 * wi] = (float)lo; o[2 * wi + 1] = (float)hi; }
                        if (FULL) {
                            const int e0 = vix * 8 + wi * 2; // element index inside the row
                            if (!CHECK) {
                                row_sum += lo + hi;
                                const unsigned a = (unsigned)(lo * lo) + (unsigned)(hi * hi); // <= 2^31
                                sq_lo += a;
                                sq_hi += (sq_lo < a);
                                hist_add(hist_s32, lo, 1u);
                                hist_add(hist_s32, hi, 1u);
                            } else {
                                if (e0 < valid_e) {
                                    row_sum += lo;
                                    const unsigned a = (unsigned)(lo * lo);
                                    sq_lo += a; sq_hi += (sq_lo < a);
                                    hist_add(hist_s32, lo, 1u);
                                }
                                if (e0 + 1 < valid_e) {
                                    row_sum += hi;
                                    const unsigned a = (unsigned)(hi * hi);
                                    sq_lo += a; sq_hi += (sq_lo < a);
                                    hist_add(hist_s32, hi, 1u);
                                }
                            }
                            if (e0 == 0) first_v = lo;
                            if (e0 + 2 == G::elems) last_v = hi;
                        }
                    }
#pragma unroll
                    for (int q4 = 0; q4 < per_vec / 4; ++q4) {
                        const int j = vix * per_vec + q4 * 4;
                        const float4 hv = *reinterpret_cast<const float4 *>(hrow + j);
                        float4 r4;
                        r4.x = o[q4 * 4] * hv.x; r4.y = o[q4 * 4 + 1] * hv.y;
                        r4.z = o[q4 * 4 + 2] * hv.z; r4.w = o[q4 * 4 + 3] * hv.w;
                        *reinterpret_cast<float4 *>(fout + j) = r4;
                    }
                }
                if (FULL && valid_e > 0) {
                    ts.sum += row_sum;
                    ts.sumsq += ((unsigned long long)sq_hi << 32) | sq_lo;
                    // first / last non-zero sample: the row's end samples decide unless one of them is zero
                    // or the row is cut by the song's end (then scan the row)
                    if (!CHECK && first_v != 0 && last_v != 0) {
                        ts.first_nz = min(ts.first_nz, (int)i0);
                        ts.last_nz = max(ts.last_nz, (int)i0 + G::elems - 1);
                    } else {
                        const short *rs = reinterpret_cast<const short *>(rowp);
                        for (int i = 0; i < valid_e; ++i)
                            if (rs[i] != 0) {
                                ts.first_nz = min(ts.first_nz, (int)i0 + i);
                                ts.last_nz = max(ts.last_nz, (int)i0 + i);
                            }
                    }
                }
            };
            if (!FULL || valid_e == G::elems) do_row(std::false_type{});
            else do_row(std::true_type{});
        }
        __syncthreads(); // rows consumed, fin complete

        // ---- prefetch the next tile of this CTA while the FFTs run
        if (tile + sd.n_parts < sd.n_tiles) issue_tile(tile + sd.n_parts);

        // ---- 8 x (512-point real FFT + power accumulation), 16 threads per frame; the two half-warps of
        // a warp run in lockstep (full-mask syncs), a half-warp past the song's last frame is ignored
        if (tile * kP1FramesPerTile + (grp & ~1) < sd.n_frames) {
            const unsigned full = 0xffffffffu;
            const bool frame_ok = tile * kP1FramesPerTile + grp < sd.n_frames;
            float2 *xchg = reinterpret_cast<float2 *>(fin + grp * kFinFrame);
            float2 v[16];
#pragma unroll
            for (int a = 0; a < 16; ++a) v[a] = *reinterpret_cast<const float2 *>(fin + grp * kFinFrame + 36 * a + 2 * lane16);
            __syncwarp(full);
#if BLX_P1_TW_REG
            fft256_halfwarp_regtw<float>(v, lane16, xchg, twr, full);
#else
            fft256_halfwarp<float>(v, lane16, xchg, tw1, full);
#endif
            __syncwarp(full);
#pragma unroll
            for (int r = 0; r < 16; ++r) xchg[lane16 + 16 * fft16_out_index(r)] = v[r];
            __syncwarp(full);
            if (frame_ok) {
#pragma unroll
                for (int d = 0; d < 8; ++d) {
                    const int k = lane16 + 16 * d;
                    if (k != 0) { // bins k and 512/2 - k from Z[k], Z[256 - k]
                        const float2 A = v[fft16_reg_of(d)];
                        const float2 B = xchg[256 - k];
                        const float2 wk = __ldg(tw2 + k);
                        const float sr = A.x + B.x, si = A.y - B.y;
                        const float dr = A.x - B.x, di = A.y + B.y;
                        const float tr = dr * wk.x - di * wk.y;
                        const float ti = dr * wk.y + di * wk.x;
                        const float ar = sr + ti, ai = si - tr;
                        const float br = sr - ti, bi = si + tr;
                        accA[d] += 0.25f * (ar * ar + ai * ai);
                        accB[d] += 0.25f * (br * br + bi * bi);
                    }
                }
                if (lane16 == 0) {
                    const float2 A = v[fft16_reg_of(8)]; // Z[128] -> X[128] = conj(Z[128])
                    acc128 += A.x * A.x + A.y * A.y;
                }
            }
        }
        __syncthreads(); // fin / xchg free for the next tile
    }

    // ---- CTA partial spectrum: fixed-order reduction over the 8 groups
    float *red = fin; // [8][256]
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        const int k = lane16 + 16 * d;
        red[grp * 256 + k] = accA[d];
        if (k != 0) red[grp * 256 + 256 - k] = accB[d];
    }
    if (lane16 == 0) red[grp * 256 + 128] = acc128;
    __syncthreads();
    for (int k = tid; k < 256; k += kP1Threads) {
        float s = 0.0f;
#pragma unroll
        for (int g2 = 0; g2 < kP1FramesPerTile; ++g2) s += red[g2 * 256 + k];
        p.partials[(size_t)(sd.part_off + part) * 256 + k] = (k == 0) ? 0.0f : s;
    }

    if (FULL) {
        // statistics: warp shuffle reduction, then one atomic per warp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ts.sum += __shfl_xor_sync(0xffffffffu, ts.sum, o);
            ts.sumsq += __shfl_xor_sync(0xffffffffu, ts.sumsq, o);
            ts.first_nz = min(ts.first_nz, __shfl_xor_sync(0xffffffffu, ts.first_nz, o));
            ts.last_nz = max(ts.last_nz, __shfl_xor_sync(0xffffffffu, ts.last_nz, o));
        }
        SongStats *st = p.stats + blockIdx.y;
        if ((tid & 31) == 0) {
            atomicAdd(reinterpret_cast<unsigned long long *>(&st->sum), (unsigned long long)ts.sum);
            atomicAdd(&st->sumsq, ts.sumsq);
            if (ts.last_nz >= 0) {
                atomicMax(&st->first_inv, 0x7fffffffu - (unsigned)ts.first_nz);
                atomicMax(&st->last_p1, (unsigned)ts.last_nz + 1u);
            }
        }
        __syncthreads();
        unsigned *ghist = p.hist + (size_t)blockIdx.y * kHistStride;
        for (int i = tid; i < kHistBins; i += kP1Threads) {
            const unsigned c = hist[i];
            if (c) atomicAdd(&ghist[i], c);
        }
    }
}

// ---------------------------------------------------------------- tensor maps (float32 input)
cudaError_t make_pass1_maps(Pass1Params *p, const void *d_pcm, long long rows) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    if (rows < 1) rows = 1;
    if ((reinterpret_cast<uintptr_t>(d_pcm) & 15) || rows > 0x7fffffffll) return cudaErrorInvalidValue;
    const cuuint64_t dims[2] = {32, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {256};
    const cuuint32_t box[2] = {32, (cuuint32_t)P1Smem<kInF32>::f32_rows};
    const cuuint32_t estr[2] = {1, 1};
    unsigned char *base = const_cast<unsigned char *>(static_cast<const unsigned char *>(d_pcm));
    for (int h = 0; h < 2; ++h) {
        const CUresult r = encode(h ? &p->map_b : &p->map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base + 128 * h, dims, strides, box,
                                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
    return cudaSuccess;
}

// ---------------------------------------------------------------- launcher
template <int KIND, bool FULL> static cudaError_t launch_one(const Pass1Params &p, int max_parts, int n_songs, cudaStream_t st) {
    using SM = P1Smem<KIND>;
    const int bytes = FULL ? SM::bytes_full : SM::bytes_lite;
    static PerDeviceOnce once;
    const cudaError_t e0 =
        once.run([bytes] { return cudaFuncSetAttribute(pass1_kernel<KIND, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); });
    if (e0 != cudaSuccess) return e0;
    dim3 grid((unsigned)max_parts, (unsigned)n_songs);
    pass1_kernel<KIND, FULL><<<grid, kP1Threads, bytes, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_pass1(int kind, bool full, const Pass1Params &p, int max_parts, int n_songs, cudaStream_t st) {
    switch (kind) {
        case kInF32: return full ? launch_one<kInF32, true>(p, max_parts, n_songs, st) : launch_one<kInF32, false>(p, max_parts, n_songs, st);
        case kInS16Stereo: return full ? launch_one<kInS16Stereo, true>(p, max_parts, n_songs, st) : launch_one<kInS16Stereo, false>(p, max_parts, n_songs, st);
        case kInS16Mono: return full ? launch_one<kInS16Mono, true>(p, max_parts, n_songs, st) : launch_one<kInS16Mono, false>(p, max_parts, n_songs, st);
        default: return cudaErrorInvalidValue;
    }
}

int pass1_tile_msamples() { return kP1TileM; }

} // namespace blx
