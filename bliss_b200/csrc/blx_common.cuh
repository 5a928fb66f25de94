// blx_common.cuh — shared definitions of the B200 bliss engine kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "../../include/blx.h"
#include "../../include/blx_frontend.h"

namespace blx {

// ---------------------------------------------------------------- constants
constexpr int kWin = 512;            // FFT length of both analysers (reference src/frequency_sort.c:6-8,
                                     // src/tempo_atk_sort.c:50)
constexpr int kHop = 256;            // envelope hop (reference src/tempo_atk_sort.c:55,120)
constexpr int kHistLo = 30864;       // first histogram bin that can influence the integral:
constexpr int kHistHi = 34670;       //   31767 - 3*301 .. 33767 + 3*301
constexpr int kHistBins = kHistHi - kHistLo + 1; // 3807
constexpr int kHistStride = 3808;
constexpr int kIntLo = 31767, kIntHi = 33767;    // reference src/amplitude_sort.c:9-10,69
constexpr int kSmoothPasses = 301;   // for (g = 0; g <= N_PASSES; ++g), reference src/amplitude_sort.c:41

// Input kinds of the streaming kernels
constexpr int kInS16Stereo = 0; // int16 interleaved L,R
constexpr int kInS16Mono = 1;   // int16 mono (channels == 1, reference src/frequency_sort.c:76-80)
constexpr int kInF32 = 2;       // float32 mono 44.1 kHz through the front-end (blx_frontend.h)

// Per-song descriptor, device resident (built on the host by the engine).
struct SongDesc {
    long long pcm_off;   // element offset of the song in the packed input buffer
    long long q_off;     // element offset of the song's decimated int16 stream (F32 input only)
    long long env_off;   // offset (doubles) of the song's envelope row t1[2F]
    int n_elems;         // input elements (int16 values, or float samples for F32)
    int n_samples;       // nSamples as the reference counts it (int16 values over all channels)
    int n_msamples;      // per-channel frames feeding the frequency analyser (n_samples / channels)
    int n_frames;        // floor(n_msamples / 512)              reference src/frequency_sort.c:50
    int n_tiles;         // pass-1 tiles covering n_msamples
    int F;               // floor(n_samples / 512)               reference src/tempo_atk_sort.c:63-64
    int n_hops;          // 2F - 2                               reference src/tempo_atk_sort.c:66-67,120
    unsigned duration;   // whole seconds                        reference src/decode.c:235
    int kind;            // kInS16Stereo / kInS16Mono / kInF32
    int part_off;        // first partial-spectrum slot of this song
    int n_parts;         // number of pass-1 CTAs (partials) for this song
    int pad0;
};

// Integer statistics gathered by pass 1 (exact, order independent).
struct SongStats {
    long long sum;          // sum of all samples                (bl_mean,     reference src/helpers.c:30-37)
    unsigned long long sumsq; // sum of squares                  (bl_variance, reference src/helpers.c:39-49)
    // Both bounds are kept in a form whose neutral element is 0 and whose merge is atomicMax, so a
    // plain memset initialises the record:
    unsigned first_inv;     // 0x7fffffff - (first index with a non-zero sample); 0 = none
                            //   (reference src/amplitude_sort.c:26-28)
    unsigned last_p1;       // (last index with a non-zero sample) + 1; 0 = none
                            //   (reference src/amplitude_sort.c:29-31)
};

// Per-song values handed from the epilogue to the envelope kernels.
struct SongNorm {
    double mean_d;     // mean / 32768                            reference src/tempo_atk_sort.c:105
    double inv_var_d;  // 1 / (variance / 32768 / 32768)          reference src/tempo_atk_sort.c:106-107
    float amplitude;
    float frequency;
    int status;
    int mean;          // bl_mean      reference src/helpers.c:30-37
    int variance;      // bl_variance  reference src/helpers.c:39-49
    int pad;
};

// ---------------------------------------------------------------- mbarrier / 1-D TMA (cp.async.bulk)
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- host: once-per-device set-up
// cudaFuncSetAttribute applies to the current device only; a process may run several engines per GPU, from
// several threads (blx.h). run(f) calls f exactly once per device, under a lock, and every caller returns only
// after it has completed.
struct PerDeviceOnce {
    std::mutex mu;
    bool done[64] = {false};
    template <typename F> cudaError_t run(F f) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64) return f();
        std::lock_guard<std::mutex> lock(mu);
        if (done[dev]) return cudaSuccess;
        e = f();
        if (e == cudaSuccess) done[dev] = true;
        return e;
    }
};

// ---------------------------------------------------------------- small helpers
__device__ __forceinline__ double int_to_double_exact(int v) {
    // 2^52 + 2^31 + v is exactly representable; one integer op + one DADD instead of I2F.F64
    return __hiloint2double(0x43300000, (int)((unsigned)v ^ 0x80000000u)) - 4503601774854144.0;
}

} // namespace blx
