// peaks.cu — the roofline denominators this engine's kernels are judged against, measured on the device at hand
// (kernel id none; not part of the analysis path). MEASURED_PEAKS.json carries the HBM copy rate and the bf16
// tensor rate only; the envelope kernel is bound by the FP64 pipe, so bench.py measures that peak itself, in the
// same process and under the same clocks as the step it is compared with (blx_measure_fp64_peak).
#include "blx_common.cuh"
#include "kernels.h"

namespace blx {

namespace {
// 8 independent FMA chains per thread: enough to cover the ~13-cycle DFMA latency at 2 cycles per issue
__global__ void __launch_bounds__(1024) dfma_peak_kernel(double *out, int iters) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 0.5;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
} // namespace

// Launches the DFMA kernel once on `st`; flops = 2 * 8 * iters * threads.
cudaError_t launch_dfma_peak(double *d_scratch, int blocks, int threads, int iters, cudaStream_t st) {
    dfma_peak_kernel<<<blocks, threads, 0, st>>>(d_scratch, iters);
    return cudaGetLastError();
}

} // namespace blx
