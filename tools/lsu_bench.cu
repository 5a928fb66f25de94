// lsu_bench.cu — how many L1/shared data-pipe cycles the access shapes of envelope_kernel cost on B200:
// 128-bit / 64-bit shared loads with 32 distinct addresses or with the two half-warps reading the same 16,
// the same through L1 (LDG), and warp shuffles. Prints cycles per warp instruction per SM with 16 warps resident.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/lsu_bench tools/lsu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096, kUnroll = 8;

template <int MODE> __global__ void __launch_bounds__(512, 1) k(const double2 *g, double *out, long long *cyc) {
    __shared__ __align__(16) double2 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = g[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int idx_dist = lane, idx_dup = lane & 15;
    double acc = 0;
    int off = warp * 32;
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int o = (off + 64 * u) & 511;
            if (MODE == 0) { const double2 t = sm[o + idx_dist]; acc += t.x + t.y; }
            if (MODE == 1) { const double2 t = sm[o + idx_dup]; acc += t.x + t.y; }
            if (MODE == 2) { const double t = reinterpret_cast<const double *>(sm)[o + idx_dist]; acc += t; }
            if (MODE == 3) { const double t = reinterpret_cast<const double *>(sm)[o + idx_dup]; acc += t; }
            if (MODE == 4) { const double2 t = __ldg(g + o + idx_dist); acc += t.x + t.y; }
            if (MODE == 5) { const double2 t = __ldg(g + o + idx_dup); acc += t.x + t.y; }
            if (MODE == 6) { acc += __shfl_xor_sync(0xffffffffu, acc, 1 + u); }          // 2 SHFL.32
            if (MODE == 7) { sm[o + idx_dist] = make_double2(acc, acc); }                  // STS.128
            if (MODE == 8) { reinterpret_cast<double *>(sm)[o + idx_dist] = acc; }         // STS.64
            if (MODE == 9) { const float t = reinterpret_cast<const float *>(sm)[o + idx_dist]; acc += t; } // LDS.32
        }
        off += 32;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char *name, const double2 *g, double *out, long long *cyc, int per_iter) {
    k<MODE><<<148, 512>>>(g, out, cyc);
    k<MODE><<<148, 512>>>(g, out, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < 148; ++i) s += (double)h[i];
    s /= 148;
    const double insts = 16.0 * kIters * kUnroll * per_iter; // warp instructions per SM
    printf("%-44s %6.2f cycles per warp instruction per SM\n", name, s / insts);
}

int main() {
    double2 *g; double *out; long long *cyc;
    cudaMalloc(&g, 1024 * sizeof(double2)); cudaMemset(g, 0, 1024 * sizeof(double2));
    cudaMalloc(&out, 148 * 512 * sizeof(double)); cudaMalloc(&cyc, 148 * sizeof(long long));
    run<0>("LDS.128, 32 distinct", g, out, cyc, 1);
    run<1>("LDS.128, half-warps read the same 16", g, out, cyc, 1);
    run<2>("LDS.64, 32 distinct", g, out, cyc, 1);
    run<3>("LDS.64, half-warps read the same 16", g, out, cyc, 1);
    run<4>("LDG.128 (L1 hit), 32 distinct", g, out, cyc, 1);
    run<5>("LDG.128 (L1 hit), half-warps read the same 16", g, out, cyc, 1);
    run<6>("SHFL.BFLY 32-bit (x2 per double)", g, out, cyc, 2);
    run<7>("STS.128, 32 distinct", g, out, cyc, 1);
    run<8>("STS.64, 32 distinct", g, out, cyc, 1);
    run<9>("LDS.32, 32 distinct", g, out, cyc, 1);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
