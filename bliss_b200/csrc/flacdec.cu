// flacdec.cu — FLAC frames decoded on the device (row N1: the decode stage in front of the analysers).
//
// The frames of a FLAC stream are independent; inside a frame everything is sequential (the Rice-coded residual is a
// variable-length bit stream, LPC synthesis a recurrence, and a channel's subframe starts where the previous one ends).
// So: ONE WORKING THREAD PER FRAME (alone in its warp, see the kernel), thousands of frames per song. The host finds the chain of frames (header CRC-8, coded frame
// numbers: flac_reader.c, decode_flac_parallel) and uploads the file as it is; every thread runs the same frame decoder
// as the host (host/flac_core.h, compiled for the device), checks the frame's CRC-16 and that it ends where the next
// frame begins, and writes its samples - interleaved, int16 for 16-bit streams - where they belong. Any frame that
// does not check out raises one flag and the caller falls back to the resynchronising host decoder.
// Lanes of a warp walk different frames with the same instruction stream (sample loop, one-step Rice decode, LPC dot
// product); they diverge only on partition boundaries, escapes and different predictor orders.
#include "blx_common.cuh"
#include "kernels.h"

#include <cstdlib>
#include <vector>

// The frame decoder is compiled for the device AND for the host: the host instance of exactly this code is what
// blx_flac_decode_frames_emulated runs, so the device path's logic (planar scratch, plain LPC loop, byte-wise CRC) is
// testable without a GPU (tests/test_flac_reader.py).
#define FLAC_FN __host__ __device__ static inline
namespace {
__device__ unsigned short g_crc16_tab[256];
unsigned short h_crc16_tab[256];
__host__ __device__ static inline unsigned short crc16_entry(int i) {
    unsigned short c = (unsigned short)(i << 8);
    for (int k = 0; k < 8; ++k) c = (unsigned short)((c & 0x8000) ? (c << 1) ^ 0x8005 : (c << 1));
    return c;
}
__host__ __device__ static inline unsigned short dev_crc16(const unsigned char *p, size_t n) {
#ifdef __CUDA_ARCH__
    const unsigned short *tab = g_crc16_tab;
#else
    const unsigned short *tab = h_crc16_tab;
#endif
    unsigned short c = 0;
    for (size_t i = 0; i < n; ++i) c = (unsigned short)((c << 8) ^ tab[(c >> 8) ^ p[i]]);
    return c;
}
} // namespace
#define FLAC_CRC16(p, n) dev_crc16(p, n)
#include "../host/flac_core.h"

namespace blx {

namespace {
__global__ void flac_crc_init_kernel() { g_crc16_tab[threadIdx.x] = crc16_entry(threadIdx.x); }

// Frame k of the chain, part 1 (sequential): decode into its planar scratch and check it. false = the frame is bad.
__host__ __device__ static inline bool flac_frame_decode(const FlacDecodeParams &p, int k) {
    const flac_hdr h = reinterpret_cast<const flac_hdr *>(p.hdr)[k];
    int *sc = p.scratch + (size_t)p.first[k] * (size_t)p.channels; // planar: channel c at sc + c * blocksize
    size_t end = 0;
    const int rc = decode_frame(p.data, p.n_bytes, &h, sc, (size_t)h.blocksize, &end);
    const size_t want = (k + 1 < p.n_frames) ? reinterpret_cast<const flac_hdr *>(p.hdr)[k + 1].off : 0;
    return rc == 0 && !(want && end != want);
}
// part 2 (parallel over `step` workers): planar scratch -> interleaved output
__host__ __device__ static inline void flac_frame_interleave(const FlacDecodeParams &p, int k, int worker, int step) {
    const int bs = reinterpret_cast<const flac_hdr *>(p.hdr)[k].blocksize, nch = p.channels;
    const size_t base = (size_t)p.first[k] * (size_t)nch;
    const int *sc = p.scratch + base;
    const int total = bs * nch;
    if (p.out16) {
        short *o = static_cast<short *>(p.out) + base;
        for (int i = worker; i < total; i += step) o[i] = (short)sc[(size_t)(i % nch) * bs + i / nch];
    } else {
        int *o = static_cast<int *>(p.out) + base;
        for (int i = worker; i < total; i += step) o[i] = sc[(size_t)(i % nch) * bs + i / nch];
    }
}

// One WARP per frame, one working lane: the lanes of a warp cannot share a frame (everything in it is sequential) and
// must not hold different frames either - their control flow differs at every sample (refills, Rice escapes, predictor
// orders) and the warp would run at the speed of the sum of its lanes (measured: 18.4 ms for 960 frames that way).
// Alone in its warp a lane runs its frame at full speed, and a song has thousands of frames to fill the SMs with such
// warps; the idle lanes join for the interleaving pass.
__global__ void __launch_bounds__(32) flac_decode_kernel(FlacDecodeParams p) {
    const int k = blockIdx.x;
    __shared__ int ok;
    if (threadIdx.x == 0) {
        ok = flac_frame_decode(p, k) ? 1 : 0;
        if (!ok) atomicExch(p.fail, 1);
        __threadfence_block();
    }
    __syncwarp();
    if (ok) flac_frame_interleave(p, k, (int)threadIdx.x, 32);
}
} // namespace

static_assert(sizeof(flac_hdr) == 40, "flac_hdr is shared with the host reader");

void h_crc16_tab_set(int i) { h_crc16_tab[i] = crc16_entry(i); }
bool flac_decode_on_host(const FlacDecodeParams &p) {
    for (int k = 0; k < p.n_frames; ++k) {
        if (!flac_frame_decode(p, k)) return false;
        flac_frame_interleave(p, k, 0, 1);
    }
    return true;
}

cudaError_t launch_flac_decode(const FlacDecodeParams &p, cudaStream_t st) {
    static PerDeviceOnce once;
    const cudaError_t e0 = once.run([&] {
        flac_crc_init_kernel<<<1, 256, 0, st>>>();
        const cudaError_t e = cudaGetLastError();
        return e != cudaSuccess ? e : cudaStreamSynchronize(st); // other engines' streams read the table too
    });
    if (e0 != cudaSuccess) return e0;
    if (p.n_frames <= 0) return cudaSuccess;
    flac_decode_kernel<<<p.n_frames, 32, 0, st>>>(p);
    return cudaGetLastError();
}

} // namespace blx

// The same work list on the host, frame after frame, with the host instance of the device code (no GPU involved): for
// tests of the device path's logic on machines without one. All pointers are host pointers.
extern "C" int blx_flac_decode_frames_emulated(const unsigned char *file, size_t n_bytes, const void *hdr, const unsigned long long *first,
                                               int n_frames, int channels, int out16, unsigned long long samples, void *out) {
    if (!file || !hdr || !first || !out || n_frames <= 0 || channels < 1 || channels > 8 || samples == 0) return -1;
    static bool ready = false;
    if (!ready) {
        for (int i = 0; i < 256; ++i) blx::h_crc16_tab_set(i);
        ready = true;
    }
    std::vector<int> scratch((size_t)samples * channels);
    int failed = 0;
    blx::FlacDecodeParams p;
    p.data = file; p.n_bytes = n_bytes; p.hdr = hdr; p.first = first; p.n_frames = n_frames; p.channels = channels; p.out16 = out16;
    p.scratch = scratch.data(); p.out = out; p.fail = &failed;
    return blx::flac_decode_on_host(p) ? 0 : 1;
}
