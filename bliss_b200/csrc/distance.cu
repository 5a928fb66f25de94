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
// CTA = 256 threads x 2 rows each; the column vectors stream through shared memory in tiles of 2048, so
// every column is read from L2 once per 512 rows and from shared memory as a warp-wide broadcast.
// The nearest neighbour is decided on the squared distance s (float, reference operation order); the
// correctly rounded sqrt is taken only for the rare candidates with s <= best s. sqrt is monotone, so a
// candidate with a larger s can never have a strictly smaller distance; candidates that tie after
// rounding keep the lowest index (columns are visited in increasing order), exactly what a scan over
// bl_distance values gives.
namespace {
constexpr int kNearRows = 2;
constexpr int kNearTile = 2048;

__device__ __forceinline__ float sqdist_pair(const float4 a, const float4 b) {
    const float d0 = __fsub_rn(a.x, b.x), d1 = __fsub_rn(a.y, b.y), d2 = __fsub_rn(a.z, b.z), d3 = __fsub_rn(a.w, b.w);
    float s = __fmul_rn(d0, d0);
    s = __fadd_rn(s, __fmul_rn(d1, d1));
    s = __fadd_rn(s, __fmul_rn(d2, d2));
    s = __fadd_rn(s, __fmul_rn(d3, d3));
    return s;
}
} // namespace

template <bool WITH_SUM>
__global__ void __launch_bounds__(kDistThreads) distance_nearest_kernel(const float4 *__restrict__ v, int n, int row0,
                                                                        int n_rows, int *__restrict__ idx_out,
                                                                        float *__restrict__ dist_out,
                                                                        double *__restrict__ sum_out) {
    __shared__ float4 cols[kNearTile];
    const float inf = __int_as_float(0x7f800000);
    int row[kNearRows];
    float4 a[kNearRows];
    float best_s[kNearRows], best_d[kNearRows];
    int best_j[kNearRows];
    double sum[kNearRows];
#pragma unroll
    for (int r = 0; r < kNearRows; ++r) {
        const int lr = (blockIdx.x * kNearRows + r) * kDistThreads + threadIdx.x; // row inside the slab
        row[r] = (lr < n_rows) ? row0 + lr : -1;
        a[r] = (row[r] >= 0) ? v[row[r]] : make_float4(0, 0, 0, 0);
        best_s[r] = inf; best_d[r] = inf; best_j[r] = -1; sum[r] = 0.0;
    }
    for (int c0 = 0; c0 < n; c0 += kNearTile) {
        const int cn = min(kNearTile, n - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < cn; i += kDistThreads) cols[i] = v[c0 + i];
        __syncthreads();
        float part[kNearRows];
#pragma unroll
        for (int r = 0; r < kNearRows; ++r) part[r] = 0.0f;
#pragma unroll 4
        for (int i = 0; i < cn; ++i) {
            const float4 b = cols[i];
#pragma unroll
            for (int r = 0; r < kNearRows; ++r) {
                const float s = sqdist_pair(a[r], b);
                if (WITH_SUM) part[r] += __fsqrt_rn(s);
                if (s <= best_s[r] && c0 + i != row[r]) {
                    const float d = __fsqrt_rn(s);
                    if (d < best_d[r]) { best_d[r] = d; best_s[r] = s; best_j[r] = c0 + i; }
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
        if (idx_out) idx_out[o] = best_j[r];
        if (dist_out) dist_out[o] = best_d[r];
        if (WITH_SUM) sum_out[o] = sum[r];
    }
}

cudaError_t launch_distance_rows(const float *d_vectors, int n, int row0, int n_rows, int mode, float *d_out,
                                 cudaStream_t st) {
    if (n <= 0 || n_rows <= 0) return cudaSuccess;
    dim3 grid((unsigned)((n + kDistTile - 1) / kDistTile), (unsigned)((n_rows + kDistTile - 1) / kDistTile));
    distance_rows_kernel<<<grid, kDistThreads, 0, st>>>(reinterpret_cast<const float4 *>(d_vectors), n, row0, n_rows, mode,
                                                        d_out);
    return cudaGetLastError();
}

cudaError_t launch_distance_nearest(const float *d_vectors, int n, int row0, int n_rows, int *d_idx, float *d_dist,
                                    double *d_sum, cudaStream_t st) {
    if (n <= 0 || n_rows <= 0) return cudaSuccess;
    const int rows_per_cta = kDistThreads * kNearRows;
    const unsigned grid = (unsigned)((n_rows + rows_per_cta - 1) / rows_per_cta);
    const float4 *v = reinterpret_cast<const float4 *>(d_vectors);
    if (d_sum) distance_nearest_kernel<true><<<grid, kDistThreads, 0, st>>>(v, n, row0, n_rows, d_idx, d_dist, d_sum);
    else distance_nearest_kernel<false><<<grid, kDistThreads, 0, st>>>(v, n, row0, n_rows, d_idx, d_dist, d_sum);
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
