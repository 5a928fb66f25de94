// distance.cu — all-pairs distances over force vectors (kernel id BLX_K_DISTANCE) and the
// stand-alone front-end kernel.
//
// bl_distance (reference src/analyze.c:88-103): the four differences, their squares and the
// three additions are float operations evaluated left to right (tempo, amplitude, frequency,
// attack) without FMA; the sqrt is taken in double and rounded to float, which equals a
// correctly rounded float sqrt. bl_cosine_similarity (reference src/analyze.c:127-145): float
// dot product and squared norms, double sqrt / multiply / divide, rounded to float.
// Every operation below is an explicit round-to-nearest intrinsic, so the matrices are
// bit-identical to the reference's scalar code.
#include <algorithm>

#include "blx_common.cuh"
#include "kernels.h"

namespace blx {

namespace {
constexpr int kDistTile = 128; // rows per CTA
constexpr int kDistThreads = 256;

__device__ __forceinline__ float dist_pair(const float4 a, const float4 b) {
    const float d0 = __fsub_rn(a.x, b.x), d1 = __fsub_rn(a.y, b.y), d2 = __fsub_rn(a.z, b.z), d3 = __fsub_rn(a.w, b.w);
    float s = __fmul_rn(d0, d0);
    s = __fadd_rn(s, __fmul_rn(d1, d1));
    s = __fadd_rn(s, __fmul_rn(d2, d2));
    s = __fadd_rn(s, __fmul_rn(d3, d3));
    return __fsqrt_rn(s);
}

__device__ __forceinline__ float sqnorm(const float4 a) {
    float s = __fmul_rn(a.x, a.x);
    s = __fadd_rn(s, __fmul_rn(a.y, a.y));
    s = __fadd_rn(s, __fmul_rn(a.z, a.z));
    s = __fadd_rn(s, __fmul_rn(a.w, a.w));
    return s;
}

__device__ __forceinline__ float cos_pair(const float4 a, const float4 b) {
    float dot = __fmul_rn(a.x, b.x);
    dot = __fadd_rn(dot, __fmul_rn(a.y, b.y));
    dot = __fadd_rn(dot, __fmul_rn(a.z, b.z));
    dot = __fadd_rn(dot, __fmul_rn(a.w, b.w));
    const double den = __dmul_rn(__dsqrt_rn((double)sqnorm(a)), __dsqrt_rn((double)sqnorm(b)));
    return __double2float_rn(__ddiv_rn((double)dot, den));
}
} // namespace

// Materialised slab: out[(i - row0) * n + j], i in [row0, row0 + n_rows), j in [0, n).
// CTA = 128 rows x 128 columns; a thread owns one column and loops over the rows, so every warp
// store is a full 128-byte line.
__global__ void __launch_bounds__(kDistThreads) distance_rows_kernel(const float4 *__restrict__ v, int n, int row0,
                                                                     int n_rows, int mode, float *__restrict__ out) {
    __shared__ float4 rows[kDistTile];
    const int r_base = blockIdx.y * kDistTile;
    const int c_base = blockIdx.x * kDistTile;
    for (int i = threadIdx.x; i < kDistTile; i += kDistThreads) {
        const int r = r_base + i;
        rows[i] = (r < n_rows) ? v[row0 + r] : make_float4(0, 0, 0, 0);
    }
    __syncthreads();
    const int col = c_base + (threadIdx.x & (kDistTile - 1));
    const int half = threadIdx.x >> 7; // two row phases per CTA
    if (col >= n) return;
    const float4 b = v[col];
    const int r_end = min(kDistTile, n_rows - r_base);
    for (int i = half; i < r_end; i += 2) {
        const float d = (mode == 0) ? dist_pair(rows[i], b) : cos_pair(rows[i], b);
        out[(size_t)(r_base + i) * n + col] = d;
    }
}

// Fused epilogue: nearest other song (and optionally the row sum), nothing materialised.
// CTA = 256 threads x 4 rows each; the column vectors stream through shared memory in tiles of 1024, every
// component stored twice side by side, so that one 64-bit register pair holds (b.x, b.x): the squared
// distances of two rows to one column are then computed with the packed single-precision instructions of
// sm_100 for the differences and the sums (sub/add.rn.f32x2 -> FADD2: two IEEE round-to-nearest operations per
// instruction; the squares are scalar FMULs, see sq2), so each lane is bit-identical to the scalar reference sequence. Every column is read from L2 once
// per 1024 rows and from shared memory as a warp-wide broadcast.
// The nearest neighbour is decided on the squared distance s (float, reference operation order); the
// correctly rounded sqrt is taken only for the rare candidates with s <= best s. sqrt is monotone, so a
// candidate with a larger s can never have a strictly smaller distance; candidates that tie after
// rounding keep the lowest index (columns are visited in increasing order), exactly what a scan over
// bl_distance values gives.
namespace {
constexpr int kNearRows = 4;
constexpr int kNearTile = 1024;
typedef unsigned long long f32x2; // (lo, hi) = two floats

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// The squares stay scalar: ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 - which rounds
// once instead of twice - whatever --fmad says and however the two are spelled (it also folds
// fma(a, b, -0) / fma(a, 1, b) back first); a scalar mul.rn.f32 is never contracted.
__device__ __forceinline__ f32x2 sq2(f32x2 d) {
    float lo, hi;
    unpack2(d, lo, hi);
    return pack2(__fmul_rn(lo, lo), __fmul_rn(hi, hi));
}

// squared distances of two rows (packed per component in ax..aw) to the column (bx..bw, each component twice)
__device__ __forceinline__ f32x2 sqdist2(f32x2 ax, f32x2 ay, f32x2 az, f32x2 aw, f32x2 bx, f32x2 by, f32x2 bz, f32x2 bw) {
    const f32x2 d0 = sub2(ax, bx), d1 = sub2(ay, by), d2 = sub2(az, bz), d3 = sub2(aw, bw);
    f32x2 s = sq2(d0);
    s = add2(s, sq2(d1));
    s = add2(s, sq2(d2));
    s = add2(s, sq2(d3));
    return s;
}
} // namespace

// gridDim.y > 1 splits the columns into contiguous ranges of `cols_per_split` (a multiple of the tile), so that
// a slab of few rows still fills the GPU (multi-GPU runs hand every rank n / world rows): the per-range
// winners are then merged with one 64-bit atomicMin per row on (distance bits, index) - distances are
// non-negative, so the packed order is "smaller distance, then lower index", the rule of the scan itself -
// and nearest_unpack_kernel writes the two arrays. Row sums are only produced unsplit.
template <bool WITH_SUM>
__global__ void __launch_bounds__(kDistThreads) distance_nearest_kernel(const float4 *__restrict__ v, int n, int row0,
                                                                        int n_rows, int *__restrict__ idx_out,
                                                                        float *__restrict__ dist_out,
                                                                        double *__restrict__ sum_out, int cols_per_split,
                                                                        unsigned long long *__restrict__ packed) {
    __shared__ __align__(16) f32x2 cols[kNearTile][4]; // (x,x) (y,y) (z,z) (w,w)
    const float inf = __int_as_float(0x7f800000);
    int row[kNearRows];
    f32x2 ax[kNearRows / 2], ay[kNearRows / 2], az[kNearRows / 2], aw[kNearRows / 2];
    float best_s[kNearRows], best_d[kNearRows];
    int best_j[kNearRows];
    double sum[kNearRows];
    {
        float4 a[kNearRows];
#pragma unroll
        for (int r = 0; r < kNearRows; ++r) {
            const int lr = (blockIdx.x * kNearRows + r) * kDistThreads + threadIdx.x; // row inside the slab
            row[r] = (lr < n_rows) ? row0 + lr : -1;
            a[r] = (row[r] >= 0) ? v[row[r]] : make_float4(0, 0, 0, 0);
            best_s[r] = inf; best_d[r] = inf; best_j[r] = -1; sum[r] = 0.0;
        }
#pragma unroll
        for (int h = 0; h < kNearRows / 2; ++h) {
            ax[h] = pack2(a[2 * h].x, a[2 * h + 1].x); ay[h] = pack2(a[2 * h].y, a[2 * h + 1].y);
            az[h] = pack2(a[2 * h].z, a[2 * h + 1].z); aw[h] = pack2(a[2 * h].w, a[2 * h + 1].w);
        }
    }
    const int c_begin = blockIdx.y * cols_per_split, c_end = min(n, c_begin + cols_per_split);
    for (int c0 = c_begin; c0 < c_end; c0 += kNearTile) {
        const int cn = min(kNearTile, c_end - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < cn; i += kDistThreads) {
            const float4 b = v[c0 + i];
            cols[i][0] = pack2(b.x, b.x); cols[i][1] = pack2(b.y, b.y);
            cols[i][2] = pack2(b.z, b.z); cols[i][3] = pack2(b.w, b.w);
        }
        __syncthreads();
        float part[kNearRows];
#pragma unroll
        for (int r = 0; r < kNearRows; ++r) part[r] = 0.0f;
#pragma unroll 4
        for (int i = 0; i < cn; ++i) {
            const ulonglong2 b01 = *reinterpret_cast<const ulonglong2 *>(&cols[i][0]);
            const ulonglong2 b23 = *reinterpret_cast<const ulonglong2 *>(&cols[i][2]);
            float sq[kNearRows];
#pragma unroll
            for (int h = 0; h < kNearRows / 2; ++h)
                unpack2(sqdist2(ax[h], ay[h], az[h], aw[h], b01.x, b01.y, b23.x, b23.y), sq[2 * h], sq[2 * h + 1]);
            if (WITH_SUM) {
#pragma unroll
                for (int r = 0; r < kNearRows; ++r) part[r] += __fsqrt_rn(sq[r]);
            }
            // one (rarely taken) branch per column for all rows of the thread; the row's own column is sorted
            // out inside (it always gets here: its s is 0)
            bool cand = false;
#pragma unroll
            for (int r = 0; r < kNearRows; ++r) cand = cand | (sq[r] <= best_s[r]);
            if (cand) {
#pragma unroll
                for (int r = 0; r < kNearRows; ++r) {
                    if (sq[r] <= best_s[r] && c0 + i != row[r]) {
                        const float d = __fsqrt_rn(sq[r]);
                        if (d < best_d[r]) { best_d[r] = d; best_s[r] = sq[r]; best_j[r] = c0 + i; }
                    }
                }
            }
        }
        if (WITH_SUM) {
#pragma unroll
            for (int r = 0; r < kNearRows; ++r) sum[r] += (double)part[r];
        }
    }
#pragma unroll
    for (int r = 0; r < kNearRows; ++r) {
        if (row[r] < 0) continue;
        const int o = row[r] - row0;
        if (packed) {
            atomicMin(&packed[o], ((unsigned long long)__float_as_uint(best_d[r]) << 32) | (unsigned)best_j[r]);
            continue;
        }
        if (idx_out) idx_out[o] = best_j[r];
        if (dist_out) dist_out[o] = best_d[r];
        if (WITH_SUM) sum_out[o] = sum[r];
    }
}

__global__ void nearest_unpack_kernel(const unsigned long long *__restrict__ packed, int n_rows, int *__restrict__ idx_out,
                                      float *__restrict__ dist_out) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_rows) return;
    const unsigned long long p = packed[o];
    if (idx_out) idx_out[o] = (int)(unsigned)p;
    if (dist_out) dist_out[o] = __uint_as_float((unsigned)(p >> 32));
}

cudaError_t launch_distance_rows(const float *d_vectors, int n, int row0, int n_rows, int mode, float *d_out,
                                 cudaStream_t st) {
    if (n <= 0 || n_rows <= 0) return cudaSuccess;
    dim3 grid((unsigned)((n + kDistTile - 1) / kDistTile), (unsigned)((n_rows + kDistTile - 1) / kDistTile));
    distance_rows_kernel<<<grid, kDistThreads, 0, st>>>(reinterpret_cast<const float4 *>(d_vectors), n, row0, n_rows, mode,
                                                        d_out);
    return cudaGetLastError();
}

int distance_nearest_splits(int n, int n_rows, bool with_sum) {
    if (with_sum || n <= 0 || n_rows <= 0) return 1;
    const int rows_per_cta = kDistThreads * kNearRows;
    const int row_blocks = (n_rows + rows_per_cta - 1) / rows_per_cta;
    int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int want = (4 * sms + row_blocks - 1) / row_blocks;       // ~4 CTAs per SM
    const int tiles = (n + kNearTile - 1) / kNearTile;
    return std::max(1, std::min(want, std::min(tiles, 64)));
}

cudaError_t launch_distance_nearest(const float *d_vectors, int n, int row0, int n_rows, int *d_idx, float *d_dist,
                                    double *d_sum, unsigned long long *d_packed, int splits, cudaStream_t st) {
    if (n <= 0 || n_rows <= 0) return cudaSuccess;
    const int rows_per_cta = kDistThreads * kNearRows;
    const unsigned row_blocks = (unsigned)((n_rows + rows_per_cta - 1) / rows_per_cta);
    const float4 *v = reinterpret_cast<const float4 *>(d_vectors);
    if (d_sum || splits <= 1 || !d_packed) {
        if (d_sum) distance_nearest_kernel<true><<<row_blocks, kDistThreads, 0, st>>>(v, n, row0, n_rows, d_idx, d_dist, d_sum, n, nullptr);
        else distance_nearest_kernel<false><<<row_blocks, kDistThreads, 0, st>>>(v, n, row0, n_rows, d_idx, d_dist, d_sum, n, nullptr);
        return cudaGetLastError();
    }
    const int tiles = (n + kNearTile - 1) / kNearTile;
    const int cols_per_split = (tiles + splits - 1) / splits * kNearTile;
    const unsigned gy = (unsigned)((n + cols_per_split - 1) / cols_per_split);
    cudaError_t e = cudaMemsetAsync(d_packed, 0xFF, (size_t)n_rows * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    distance_nearest_kernel<false><<<dim3(row_blocks, gy), kDistThreads, 0, st>>>(v, n, row0, n_rows, nullptr, nullptr, nullptr,
                                                                                   cols_per_split, d_packed);
    nearest_unpack_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(d_packed, n_rows, d_idx, d_dist);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- fused epilogue, cosine similarity
// Most similar other song under bl_cosine_similarity (reference src/analyze.c:127-145), nothing materialised:
// per row the column with the largest similarity, the lowest index among equal floats - what a scan over the
// reference's values gives. The exact value needs two double square roots' product and a double division per pair;
// a float estimate (dot * 1/|a| * 1/|b|, within 2e-6 of the exact value) sorts out every column that cannot reach
// the row's current best, so the exact expression runs for a handful of candidates per row only.
namespace {
constexpr int kCosTile = 1024;
struct CosCol { float4 b; float rn; float pad; double sn; }; // vector, 1 / |b| (float), |b| (double)
}

__global__ void __launch_bounds__(kDistThreads) cosine_nearest_kernel(const float4 *__restrict__ v, int n, int row0, int n_rows,
                                                                      int *__restrict__ idx_out, float *__restrict__ sim_out) {
    __shared__ CosCol cols[kCosTile];
    const int lr = blockIdx.x * kDistThreads + threadIdx.x;
    const int row = (lr < n_rows) ? row0 + lr : -1;
    const float4 a = (row >= 0) ? v[row] : make_float4(1, 0, 0, 0);
    const double sa = __dsqrt_rn((double)sqnorm(a));
    const float rna = (float)(1.0 / sa);
    float best = -__int_as_float(0x7f800000);
    int best_j = -1;
    for (int c0 = 0; c0 < n; c0 += kCosTile) {
        const int cn = min(kCosTile, n - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < cn; i += kDistThreads) {
            const float4 b = v[c0 + i];
            const double sn = __dsqrt_rn((double)sqnorm(b));
            cols[i].b = b; cols[i].sn = sn; cols[i].rn = (float)(1.0 / sn); cols[i].pad = 0.0f;
        }
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < cn; ++i) {
            const float4 b = cols[i].b;
            float dot = __fmul_rn(a.x, b.x);
            dot = __fadd_rn(dot, __fmul_rn(a.y, b.y));
            dot = __fadd_rn(dot, __fmul_rn(a.z, b.z));
            dot = __fadd_rn(dot, __fmul_rn(a.w, b.w));
            const float est = dot * rna * cols[i].rn;
            // !(est + margin < best) also lets NaN estimates through (zero vectors: the exact value is NaN and never wins)
            if (!(est + 4e-6f < best) && c0 + i != row) {
                const float c = __double2float_rn(__ddiv_rn((double)dot, __dmul_rn(sa, cols[i].sn)));
                if (c > best) { best = c; best_j = c0 + i; }
            }
        }
    }
    if (row >= 0) {
        if (idx_out) idx_out[row - row0] = best_j;
        if (sim_out) sim_out[row - row0] = best;
    }
}

cudaError_t launch_cosine_nearest(const float *d_vectors, int n, int row0, int n_rows, int *d_idx, float *d_sim, cudaStream_t st) {
    if (n <= 0 || n_rows <= 0) return cudaSuccess;
    cosine_nearest_kernel<<<(unsigned)((n_rows + kDistThreads - 1) / kDistThreads), kDistThreads, 0, st>>>(
        reinterpret_cast<const float4 *>(d_vectors), n, row0, n_rows, d_idx, d_sim);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- stand-alone front-end
// blx_frontend.h, one thread per output frame. Used by blx_frontend_f32 (tests, decode of 44.1 kHz
// files); the analysis path runs the same arithmetic fused into pass 1.
__global__ void frontend_kernel(const float *__restrict__ x, long long n_in, short *__restrict__ out) {
    const float H[BLX_FE_NPAIRS] = BLX_FE_TAPS;
    const long long n_out = n_in / 2;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_out; t += (long long)gridDim.x * blockDim.x) {
        auto at = [&](long long i) { return (i >= 0 && i < n_in) ? x[i] : 0.0f; };
        float acc = __fmul_rn(BLX_FE_CENTER, at(2 * t));
#pragma unroll
        for (int k = 0; k < BLX_FE_NPAIRS; ++k)
            acc = __fmaf_rn(H[k], __fadd_rn(at(2 * t - (2 * k + 1)), at(2 * t + (2 * k + 1))), acc);
        float q = rintf(acc);
        q = fminf(fmaxf(q, -32768.0f), 32767.0f);
        const short qs = (short)(int)q;
        reinterpret_cast<short2 *>(out)[t] = make_short2(qs, qs);
    }
}

cudaError_t launch_frontend(const float *d_in, long long n_in, short *d_out, cudaStream_t st) {
    const long long n_out = n_in / 2;
    if (n_out <= 0) return cudaSuccess;
    const int threads = 256;
    const long long blocks = (n_out + threads - 1) / threads;
    frontend_kernel<<<(unsigned)(blocks > 65535 * 16 ? 65535 * 16 : blocks), threads, 0, st>>>(d_in, n_in, d_out);
    return cudaGetLastError();
}

} // namespace blx
