// kernels.h — launch interfaces between the engine (engine.cu) and the kernel translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "blx_common.cuh"

namespace blx {

struct Pass1Params {
    const void *pcm;        // packed input buffer (int16 or float)
    const SongDesc *songs;  // descriptors of the songs of this launch (blockIdx.y)
    const float *hann;      // [512]
    const float2 *tw1;      // [256] exp(-2 pi i b c / 256) at [c * 16 + b]
    const float2 *tw2;      // [128] exp(-2 pi i k / 512)
    float *partials;        // [n_parts_total][256]
    unsigned *hist;         // [n_songs][kHistStride]   (FULL)
    SongStats *stats;       // [n_songs]                (FULL)
    short *qout;            // decimated int16 stream   (FULL, F32 input)
    // F32 input: the packed buffer as rows of 64 floats, first / second 128 bytes of every row
    // (2-D tensor maps {32 floats, rows}, row stride 256 bytes, box {32, 129}, 128-byte swizzle)
    CUtensorMap map_a, map_b;
};
// Fills map_a / map_b for a float32 buffer of `rows` rows of 64 floats at d_pcm (16-byte aligned).
cudaError_t make_pass1_maps(Pass1Params *p, const void *d_pcm, long long rows);
cudaError_t launch_pass1(int kind, bool full, const Pass1Params &p, int max_parts, int n_songs, cudaStream_t st);
int pass1_tile_msamples();

struct EpilogueParams {
    const SongDesc *songs;
    const float *partials;
    const unsigned *hist;
    const SongStats *stats;
    SongNorm *norm;     // out
    float *frequency;   // optional out (spectral-only entry point), may be NULL
    double *energy;     // envelope rows (SongDesc::env_off): entries past the last hop are cleared here; may be NULL
    unsigned what;      // BLX_DO_* mask
};
cudaError_t launch_epilogue(const EpilogueParams &p, int n_songs, cudaStream_t st);

struct EnvelopeParams {
    const short *stream;     // int16 samples: the S16 input itself, or the decimated stream (dup = 1)
    const SongDesc *songs;
    const SongNorm *norm;
    const double2 *tw1;      // [256] double twiddles, same layout as the float ones
    const double2 *tw2;      // [128]
    double *energy;          // rows of 2F doubles at SongDesc::env_off: E[m]
    int dup;                 // 1: logical S[i] = stream[q_off + (i >> 1)]; 0: S[i] = stream[pcm_off + i]
    int slow_chain;          // test hook (BLX_DEBUG_SLOW_CHAIN): every hop takes the fallback accumulation path
};
cudaError_t launch_envelope(const EnvelopeParams &p, int max_hops, int n_songs, cudaStream_t st);

struct TailParams {
    const SongDesc *songs;
    const SongNorm *norm;
    const double *xlog;      // log-compressed energies (launch_logcomp), rows at SongDesc::env_off
    blx_result *out;
    unsigned what;
};
cudaError_t launch_logcomp(const double *d_energy, double *d_xlog, long long n, cudaStream_t st);
cudaError_t launch_tail(const TailParams &p, int n_songs, cudaStream_t st);

cudaError_t launch_distance_rows(const float *d_vectors, int n, int row0, int n_rows, int mode, float *d_out, cudaStream_t st);
// splits > 1 (distance_nearest_splits) needs d_packed: n_rows 64-bit words of scratch
int distance_nearest_splits(int n, int n_rows, bool with_sum);
cudaError_t launch_distance_nearest(const float *d_vectors, int n, int row0, int n_rows, int *d_idx, float *d_dist,
                                    double *d_sum, unsigned long long *d_packed, int splits, cudaStream_t st);
cudaError_t launch_cosine_nearest(const float *d_vectors, int n, int row0, int n_rows, int *d_idx, float *d_sim, cudaStream_t st);
cudaError_t launch_rect_filter(double *d_out, const double *d_in, int n, int width, cudaStream_t st);
// decode-stage resampler (include/blx_resample.h): in = the reader's int32 samples (interleaved), out = int16 stereo
struct ResampleParams {
    const int *in;
    const short *in16;      // non-null: the samples are stored as int16 (16-bit PCM files), `in` is unused
    short *out;
    const float *bank_f32;  // [P][L] (float-internal kinds)
    const short *bank_s16;  // [P][L] (BLX_RS_KIND_U8)
    long long n_in, n_out;  // frames
    int kind, bits, channels;
    int L, P, q, center;    // P == 0: same rate, format conversion only
    int mono_gain_last;     // BLX_RS_MONO_GAIN_LAST(in_rate)
};
cudaError_t launch_resample(const ResampleParams &p, cudaStream_t st);

// FLAC frames on the device (flacdec.cu): one thread per frame of the chain the host found
struct FlacDecodeParams {
    const unsigned char *data;            // the file
    size_t n_bytes;
    const void *hdr;                      // flac_hdr[n_frames] (host/flac_core.h)
    const unsigned long long *first;      // first sample (per channel) of every frame
    int n_frames, channels, out16;
    int *scratch;                         // planar per frame, samples * channels ints
    void *out;                            // interleaved int16 (out16) or int32
    int *fail;                            // set to 1 by any frame that does not check out
};
cudaError_t launch_flac_decode(const FlacDecodeParams &p, cudaStream_t st);
void h_crc16_tab_set(int i);
bool flac_decode_on_host(const FlacDecodeParams &p); // the host instance of the same code (tests)
cudaError_t launch_dfma_peak(double *d_scratch, int blocks, int threads, int iters, cudaStream_t st);
cudaError_t launch_frontend(const float *d_in, long long n_in, short *d_out, cudaStream_t st);

} // namespace blx
