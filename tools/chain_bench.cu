// Micro-benchmark of float-accumulation chain variants (envelope consumer), cycles per step.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/chain_bench tools/chain_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

struct Chain {
    double Sp, C; long long L;
    __device__ __forceinline__ void rebase(double r) {
        const int ex = (__double2hiint(r) >> 20) & 0x7ff;
        if (ex >= 1023 - 126 && ex <= 1023 + 127) {
            const int chi = (ex + 29) << 20;
            C = __hiloint2double(chi, 0); Sp = r + C; L = ((long long)chi << 32) + (1ll << 24);
        } else { C = 0.0; Sp = r; L = 0; }
    }
    __device__ __forceinline__ void add(double p) {
        const double A = Sp + p;
        if (__double_as_longlong(A) < L) Sp = A;
        else { const double s = Sp - C; rebase((double)(float)(s + p)); }
    }
    __device__ __forceinline__ double value() const { return Sp - C; }
};

// variant 0: reference sequence (cvt chain); 1: per-step check; 2: groups of G with one check
template <int VAR, int G>
__device__ __forceinline__ double run_chain(const double *Pw) {
    if (VAR == 0) {
        float s = 0.0f;
#pragma unroll 4
        for (int k = 0; k <= 256; ++k) s = (float)((double)s + Pw[k]);
        return (double)s;
    }
    Chain ch; ch.rebase((double)(float)Pw[0]);
    if (VAR == 1) {
#pragma unroll 4
        for (int k = 1; k <= 256; ++k) ch.add(Pw[k]);
    } else {
        for (int k = 1; k <= 256; k += G) {
            double p[G];
#pragma unroll
            for (int i = 0; i < G; ++i) p[i] = Pw[k + i];
            double A = ch.Sp;
#pragma unroll
            for (int i = 0; i < G; ++i) A += p[i];
            if (__double_as_longlong(A) < ch.L) ch.Sp = A;
            else {
#pragma unroll
                for (int i = 0; i < G; ++i) ch.add(p[i]);
            }
        }
    }
    return ch.value();
}

template <int VAR, int G>
__global__ void bench(const double *Pg, double *out, long long *cyc, int lanes, int busy_warps, int reps) {
    extern __shared__ double P[]; // [32][257]
    for (int i = threadIdx.x; i < 32 * 257; i += blockDim.x) P[i] = Pg[(blockIdx.x % 64) * 32 * 257 + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        long long t0 = clock64();
        double acc = 0;
        for (int r = 0; r < reps; ++r)
            if (lane < lanes) acc += run_chain<VAR, G>(P + ((lane + r) & 31) * 257);
        long long t1 = clock64();
        if (lane == 0) cyc[blockIdx.x] = t1 - t0;
        if (lane < lanes) out[blockIdx.x * 32 + lane] = acc;
    } else if (warp <= busy_warps) { // FP64 background load
        double a0 = lane, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
        for (int i = 0; i < reps * 700; ++i) { a0 = fma(a0, 1.0000001, 0.5); a1 = fma(a1, 1.0000001, 0.5); a2 = fma(a2, 1.0000001, 0.5); a3 = fma(a3, 1.0000001, 0.5); }
        if (a0 + a1 + a2 + a3 == 12345.678) out[0] = a0;
    }
}

template <int VAR, int G> void go(const char *name, const double *dP, double *dout, long long *dcyc, int lanes, int busy) {
    const int reps = 8, blocks = 148 * 2;
    cudaFuncSetAttribute(bench<VAR, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 257 * 8);
    bench<VAR, G><<<blocks, 32 * (1 + busy), 32 * 257 * 8>>>(dP, dout, dcyc, lanes, busy, reps);
    cudaDeviceSynchronize();
    long long h[296]; cudaMemcpy(h, dcyc, sizeof(h), cudaMemcpyDeviceToHost);
    double hs[32]; cudaMemcpy(hs, dout, sizeof(hs), cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < blocks; ++i) m += h[i];
    printf("%-28s lanes %2d busy_warps %d: %.1f cycles/step   (E[0]=%.9g)\n", name, lanes, busy, m / blocks / reps / 257.0, hs[0]);
}

int main() {
    const int n = 64 * 32 * 257;
    double *h = (double *)malloc(n * 8);
    srand(1);
    for (int c = 0; c < 64 * 32; ++c) {
        double scale = exp(((rand() % 1000) / 1000.0 - 0.5) * 6.0);
        for (int k = 0; k < 257; ++k) { // spectrum-like: chi-square-ish power with a tilt
            double u1 = (rand() + 1.0) / (RAND_MAX + 2.0), u2 = (rand() + 1.0) / (RAND_MAX + 2.0);
            h[c * 257 + k] = scale * (-log(u1) - log(u2)) * (1.0 + 3.0 * exp(-k / 20.0));
        }
    }
    double *dP, *dout; long long *dcyc;
    cudaMalloc(&dP, n * 8); cudaMalloc(&dout, 296 * 32 * 8); cudaMalloc(&dcyc, 296 * 8);
    cudaMemcpy(dP, h, n * 8, cudaMemcpyHostToDevice);
    for (int busy = 0; busy <= 8; busy += 8) {
        for (int lanes : {1, 4, 8, 16}) {
            go<0, 1>("reference cvt chain", dP, dout, dcyc, lanes, busy);
            go<1, 1>("per-step check", dP, dout, dcyc, lanes, busy);
            go<2, 4>("group of 4", dP, dout, dcyc, lanes, busy);
            go<2, 8>("group of 8", dP, dout, dcyc, lanes, busy);
        }
    }
    return 0;
}
