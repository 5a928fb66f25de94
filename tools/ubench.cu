// Micro-benchmarks that size the envelope kernel's design choices on the B200 at hand:
// FP64 pipe throughput, latency of the float-accumulation chain (DADD + F2F + F2F), shared-memory
// atomic rates. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_tput(double *out, int iters) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 0.5;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void ffma_tput(float *out, int iters) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float m = 1.0000001f, c = 0.5f;
    for (int i = 0; i < iters; ++i) {
        a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
        a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void chain_lat(double *out, long long *cycles, const double *p, int n, int mode) {
    // one warp, lane 0 measures
    double acc = p[0];
    float facc = (float)p[0];
    long long t0 = clock64();
    if (mode == 0) { // DADD chain
        for (int i = 0; i < n; ++i) acc = acc + p[i & 255];
    } else if (mode == 1) { // reference accumulation: (float)((double)f + p)
        for (int i = 0; i < n; ++i) facc = (float)((double)facc + p[i & 255]);
    } else if (mode == 2) { // DFMA chain
        for (int i = 0; i < n; ++i) acc = fma(acc, 1.0000001, p[i & 255]);
    } else if (mode == 3) { // cvt round trip only
        for (int i = 0; i < n; ++i) acc = (double)(float)acc * 1.0;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cycles[0] = t1 - t0; out[0] = acc + facc; }
}

__global__ void acc_tput(double *out, const double *p, int n) {
    // every lane runs the reference accumulation on its own data: throughput of the F2F path
    __shared__ double sp[256];
    sp[threadIdx.x & 255] = p[threadIdx.x & 255];
    __syncthreads();
    float f0 = threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
    for (int i = 0; i < n; ++i) {
        const double v = sp[i & 255];
        f0 = (float)((double)f0 + v); f1 = (float)((double)f1 + v);
        f2 = (float)((double)f2 + v); f3 = (float)((double)f3 + v);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = f0 + f1 + f2 + f3;
}

__global__ void atoms_tput(unsigned *out, int iters, int mode) {
    __shared__ unsigned h[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) h[i] = 0;
    __syncthreads();
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x;
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        const unsigned idx = (mode == 0) ? ((x >> 12) & 4095u) : (mode == 1 ? 7u : (threadIdx.x & 31u) * 33u);
        atomicAdd(&h[idx], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = h[7];
}

template <typename F> float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s SMs %d clock %d MHz\n", prop.name, prop.multiProcessorCount, clk_khz / 1000);
    const int sms = prop.multiProcessorCount;
    double *dout; cudaMalloc(&dout, sizeof(double) * sms * 8 * 1024);
    float *fout; cudaMalloc(&fout, sizeof(float) * sms * 8 * 1024);
    double hp[256]; for (int i = 0; i < 256; ++i) hp[i] = 1.0 + i * 1e-3;
    double *dp; cudaMalloc(&dp, sizeof(hp)); cudaMemcpy(dp, hp, sizeof(hp), cudaMemcpyHostToDevice);
    long long *dcyc; cudaMalloc(&dcyc, 8);
    {
        const int iters = 20000;
        for (int tpb : {256, 512, 1024}) {
            float ms = time_ms([&] { dfma_tput<<<sms * 2, tpb>>>(dout, iters); });
            double flops = 2.0 * 8 * iters * (double)sms * 2 * tpb;
            printf("DFMA  tpb %4d: %.2f TFLOP/s  (%.1f DFMA/clk/SM at %d MHz nominal)\n", tpb, flops / ms / 1e9,
                   flops / 2 / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000);
        }
        float ms = time_ms([&] { ffma_tput<<<sms * 2, 1024>>>(fout, iters); });
        double flops = 2.0 * 8 * iters * (double)sms * 2 * 1024;
        printf("FFMA  tpb 1024: %.2f TFLOP/s\n", flops / ms / 1e9);
    }
    for (int mode = 0; mode < 4; ++mode) {
        const int n = 100000;
        chain_lat<<<1, 32>>>(dout, dcyc, dp, n, mode);
        cudaDeviceSynchronize();
        chain_lat<<<1, 32>>>(dout, dcyc, dp, n, mode);
        long long cyc; cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
        const char *names[4] = {"DADD chain", "float-accumulate chain (cvt+DADD+cvt)", "DFMA chain", "cvt round trip + DMUL"};
        printf("latency %-40s: %.1f cycles/step\n", names[mode], (double)cyc / n);
    }
    {
        const int n = 20000;
        float ms = time_ms([&] { acc_tput<<<sms * 2, 1024>>>(dout, dp, n); });
        double steps = 4.0 * n * (double)sms * 2 * 1024;
        printf("float-accumulate throughput: %.2f steps/clk/SM (nominal clock)\n", steps / (ms * 1e-3) / sms / (clk_khz * 1e3));
    }
    unsigned *uout; cudaMalloc(&uout, 4 * sms * 4);
    for (int mode = 0; mode < 3; ++mode) {
        const int iters = 20000;
        float ms = time_ms([&] { atoms_tput<<<sms * 2, 256>>>(uout, iters, mode); });
        double ops = (double)iters * sms * 2 * 256;
        const char *names[3] = {"random 4096 bins", "single address", "conflict-free"};
        printf("ATOMS %-18s: %.2f lane-atomics/clk/SM (nominal clock)\n", names[mode], ops / (ms * 1e-3) / sms / (clk_khz * 1e3));
    }
    return 0;
}
