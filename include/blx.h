/* blx.h — the thin device C-ABI of the B200-native bliss engine.
 *
 * extern "C", plain pointers and sizes, no CUDA / torch types. The 15 functions of
 * bliss.h are host-C wrappers over these entry points (the C files under bliss_b200/host); batch
 * users (bench.py, the Python package, a cgo/JNI/ctypes binding) call them directly.
 * Each entry point names the reference interface it replaces.
 *
 * All analysis work runs in hand-written sm_100a CUDA kernels (bliss_b200/csrc).
 * There is NO CPU implementation behind these calls: without a CUDA device
 * blx_init fails with BLX_ERR_CUDA and every other call with BLX_ERR_ARG.
 *
 * Threading: an engine may be used from one thread at a time; create one engine per
 * thread/stream for concurrency. blx_last_error() is thread-local.
 */
#ifndef BLX_H_
#define BLX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLX_OK 0
#define BLX_ERR_CUDA -1  /* CUDA runtime / device error (see blx_last_error) */
#define BLX_ERR_ARG -2   /* invalid argument */
#define BLX_ERR_NOMEM -3 /* host or device allocation failed */

/* Which analysers to run (bit mask). */
#define BLX_DO_AMPLITUDE 0x1u /* replaces bl_amplitude_sort, reference src/amplitude_sort.c:12-80 */
#define BLX_DO_FREQUENCY 0x2u /* replaces bl_frequency_sort, reference src/frequency_sort.c:20-140 */
#define BLX_DO_ENVELOPE 0x4u  /* replaces bl_envelope_sort,  reference src/tempo_atk_sort.c:42-296 */
#define BLX_DO_ALL 0x7u       /* + rating of reference src/analyze.c:63-79 */

/* Per-song status bits in blx_result.status (0 = analysed). The reference has no error
 * channel for these inputs: it loops forever, divides by zero or reads out of bounds
 * (SURVEY.md §7.3 H5); the engine flags them and returns NaN ratings instead. */
#define BLX_SONG_TOO_SHORT 0x1 /* nSamples < 1536 (no complete envelope hop) or duration == 0 */
#define BLX_SONG_SILENT 0x2    /* every sample is zero */
#define BLX_SONG_FLAT 0x4      /* integer variance is zero */

/* Result record, 32 bytes. The first 16 bytes are exactly struct force_vector_s
 * (reference include/bliss.h:26-31). */
typedef struct blx_result {
    float tempo;
    float amplitude;
    float frequency;
    float attack;
    float force;      /* reference src/analyze.c:68-72 */
    int calm_or_loud; /* BL_LOUD 0 / BL_CALM 1 / BL_UNKNOWN 2, reference src/analyze.c:73-79 */
    int beat;         /* onset count behind `tempo` (reference src/tempo_atk_sort.c:277-280) */
    int status;
} blx_result;

typedef struct blx_engine blx_engine;

/* PCM sample formats accepted by the device entry point. */
#define BLX_FMT_S16 0 /* int16, interleaved, 22 050 Hz: the analysers' native input */
#define BLX_FMT_F32 1 /* float32 mono 44.1 kHz, converted by the front-end of blx_frontend.h */

/* Every song of a packed device buffer must start at an element offset that is a
 * multiple of BLX_ALIGN_ELEMS and the buffer must stay readable up to the next
 * multiple of BLX_ALIGN_ELEMS past each song's end (contents there are ignored). */
#define BLX_ALIGN_ELEMS 64

/* ---- lifecycle ------------------------------------------------------------ */
int blx_device_count(void);
int blx_init(int device, blx_engine **out);
void blx_shutdown(blx_engine *e);
const char *blx_last_error(void);

/* Optional tuning, before the first analyse call. chunk_bytes: size of each of the
 * two device staging buffers used by the host-buffer entry points (default 1 GiB). */
int blx_configure(blx_engine *e, size_t chunk_bytes);

/* Songs per kernel sequence of the device-resident entry points (default 512, or $BLX_SUB_BATCH at blx_init).
 * Results do not depend on it. */
int blx_configure_sub_batch(blx_engine *e, int songs);

/* Test hooks (results must not depend on them). BLX_DEBUG_SLOW_CHAIN: the envelope kernel takes
 * its fallback path (binade-by-binade scan) for the float accumulation of every hop. */
#define BLX_DEBUG_SLOW_CHAIN 1u
int blx_debug_flags(blx_engine *e, unsigned flags);

/* ---- per-song analysis, HOST buffers (the end-to-end path) -------------------
 * Replaces the analysis part of bl_analyze (reference src/analyze.c:43-79) for a
 * batch of songs already decoded in host memory. Copies are chunked and
 * double-buffered against the kernels. `out` is a host array of n_songs records. */
int blx_analyze_batch_s16(blx_engine *e, const int16_t *const *pcm, const int *n_samples,
                          const int *channels /* NULL => 2 */, const uint64_t *duration_s, int n_songs,
                          unsigned what, blx_result *out);

/* Same for 44.1 kHz mono float32 songs (n_in = samples per song); duration is
 * n_in / 44100 whole seconds, nSamples = 2 * (n_in / 2). */
int blx_analyze_batch_f32(blx_engine *e, const float *const *pcm, const int64_t *n_in, int n_songs,
                          unsigned what, blx_result *out);

/* Float32 songs (mono or stereo interleaved, any sample rate the resampler takes) exactly as bl_analyze would treat a
 * float file: the decode-stage resampler of include/blx_resample.h on the device (FFmpeg's libswresample, bit for
 * bit), then the native int16 analysis; duration is n_frames / in_rate whole seconds. For callers that need the
 * reference's stream rather than the engine's own, cheaper 44.1 kHz front-end (blx_analyze_batch_f32). */
int blx_analyze_batch_f32_exact(blx_engine *e, const float *const *pcm, const int64_t *n_frames, int channels, int in_rate,
                                int n_songs, unsigned what, blx_result *out);

/* ---- per-song analysis, DEVICE-resident packed buffer -------------------------
 * d_pcm: device pointer (BLX_FMT_S16: int16_t*, BLX_FMT_F32: float*). offsets/lengths
 * are HOST arrays in elements; channels (S16 only, NULL => 2) and duration_s (S16 only;
 * F32 derives it) are HOST arrays. d_out: DEVICE array of n_songs blx_result.
 * stream: a cudaStream_t passed as void* (NULL = the engine's own stream; for CUDA's legacy default stream,
 * whose handle is also 0, pass cudaStreamLegacy = (void *)1). The call only enqueues work; it does not synchronise. The kernels run on the engine's own streams (sub-batches of 512
 * songs; the sequential tail of one under the envelope kernel of the next), fenced against `stream` on both
 * sides: work enqueued on `stream` before the call is seen, work enqueued after it sees the results. */
int blx_analyze_device(blx_engine *e, int fmt, const void *d_pcm, const int64_t *offsets,
                       const int64_t *lengths, const int *channels, const uint64_t *duration_s, int n_songs,
                       unsigned what, blx_result *d_out, void *stream);

/* Same, for a job of several batches: the call returns without making `stream` wait for the results, so that the
 * latency-bound tail of this batch runs under the next batch's kernels. Inputs and d_out must stay untouched
 * until blx_join(e, stream), which makes `stream` wait for everything enqueued so far. */
int blx_analyze_device_async(blx_engine *e, int fmt, const void *d_pcm, const int64_t *offsets,
                             const int64_t *lengths, const int *channels, const uint64_t *duration_s, int n_songs,
                             unsigned what, blx_result *d_out, void *stream);
int blx_join(blx_engine *e, void *stream);

/* The fused window + rFFT-512 + per-bin power + band-ratio kernel ALONE
 * (BASELINE.json configs[1]); writes only `frequency` per song to d_frequency
 * (device float[n_songs]). Same buffer contract as blx_analyze_device. */
int blx_spectral_device(blx_engine *e, int fmt, const void *d_pcm, const int64_t *offsets,
                        const int64_t *lengths, const int *channels, int n_songs, float *d_frequency,
                        void *stream);

/* ---- distances ----------------------------------------------------------------
 * All-pairs euclidean distance with the float semantics of bl_distance
 * (reference src/analyze.c:96-100: float sub/mul/add left to right, no FMA, sqrt
 * correctly rounded). vectors: n x 4 floats (tempo, amplitude, frequency, attack). */
int blx_distance_matrix(blx_engine *e, const float *vectors, int n, float *out /* n x n, host */);
int blx_cosine_matrix(blx_engine *e, const float *vectors, int n, float *out); /* reference src/analyze.c:135-142 */

/* Device slab: rows [row0, row0 + n_rows) of the n x n matrix into d_out (n_rows x n).
 * mode 0 = euclidean, 1 = cosine similarity. */
int blx_distance_rows_device(blx_engine *e, const float *d_vectors, int n, int row0, int n_rows, int mode,
                             float *d_out, void *stream);

/* Fused epilogue for matrices that cannot be materialised (1 M x 1 M): per row of the
 * slab, the nearest other song (ties to the lowest index, as a scan over bl_distance values) and,
 * if d_row_sum is not NULL, the sum of the row's distances (a checksum: float partial sums per
 * 1024-column tile, tiles added in double). Any of the three outputs may be NULL. A slab with few
 * rows is computed in column ranges merged through a scratch array owned by the engine: keep one
 * such call per engine in flight at a time (calls on one stream are ordered anyway). */
int blx_distance_nearest_device(blx_engine *e, const float *d_vectors, int n, int row0, int n_rows,
                                int *d_nearest_index, float *d_nearest_dist, double *d_row_sum, void *stream);

/* The same for bl_cosine_similarity (reference src/analyze.c:127-145): per row of the slab the MOST similar other
 * song - the largest similarity as the reference's scalar code rounds it, the lowest index among equal values.
 * Either output may be NULL. */
int blx_cosine_nearest_device(blx_engine *e, const float *d_vectors, int n, int row0, int n_rows, int *d_index,
                              float *d_similarity, void *stream);

/* ---- several GPUs from one C process ---------------------------------------------
 * BASELINE.json configs[2] / configs[4] without Python: one engine + one host thread per device. The device
 * threads take units of up to 128 consecutive songs from a shared counter (no data-path collective; a device on a
 * slower host link ends up with fewer songs; records do not depend on which device analysed them); afterwards the
 * force vectors stay resident in contiguous blocks (device r holds songs [r N / G, (r + 1) N / G)). blx_multi_nearest all-gathers them device to device (ncclAllGather over
 * NVLink; peer copies if NCCL cannot start - blx_multi_transport says which) and reduces every device's rows
 * against all columns: nearest other song per song, with the bl_distance semantics of blx_distance_nearest_device.
 * devices == NULL: the first n_devices visible devices (n_devices <= 0: all). */
typedef struct blx_multi blx_multi;
int blx_multi_init(const int *devices, int n_devices, blx_multi **out);
void blx_multi_shutdown(blx_multi *m);
int blx_multi_device_count(blx_multi *m);
const char *blx_multi_transport(blx_multi *m);
blx_engine *blx_multi_engine(blx_multi *m, int rank);
int blx_multi_songs_taken(blx_multi *m, int rank); /* songs device `rank` analysed in the last batch */
int blx_multi_analyze_batch_s16(blx_multi *m, const int16_t *const *pcm, const int *n_samples, const int *channels,
                                const uint64_t *duration_s, int n_songs, unsigned what, blx_result *out);
int blx_multi_analyze_batch_f32(blx_multi *m, const float *const *pcm, const int64_t *n_in, int n_songs, unsigned what,
                                blx_result *out);
/* Replaces the resident vectors by n host vectors (n x 4 floats), sharded the same way. */
int blx_multi_set_vectors(blx_multi *m, const float *vectors, int n);
/* All-gather + all-pairs nearest neighbour over the resident vectors; host outputs of n entries (either may be
 * NULL); gather_ms / nearest_ms (may be NULL): device time of the two phases, max over devices. */
int blx_multi_nearest(blx_multi *m, int *nearest_index, float *nearest_dist, float *gather_ms, float *nearest_ms);

/* ---- small reference helpers on the device --------------------------------------
 * bl_mean / bl_variance (reference src/helpers.c:30-49) and bl_rectangular_filter
 * (reference src/tempo_atk_sort.c:19-40), host buffers. */
/* mean_in: NULL => variance around the array's own integer mean (what the reference's callers pass,
 * reference src/tempo_atk_sort.c:101-103); otherwise around *mean_in. The device pass returns the
 * exact integer sums; sum (s - m)^2 = sum s^2 - 2 m sum s + n m^2 is evaluated in 64-bit integers. */
int blx_mean_variance_s16(blx_engine *e, const int16_t *pcm, int n_samples, const int *mean_in, int *mean,
                          int *variance);
int blx_rectangular_filter(blx_engine *e, double *out, const double *in, int n, int width);

/* Front-end alone (blx_frontend.h): host float32 in, host int16 stereo out
 * (2 * (n_in / 2) values). */
int blx_frontend_f32(blx_engine *e, const float *pcm, int64_t n_in, int16_t *out);

/* Decode-stage resampler (include/blx_resample.h), replacing the libswresample calls of reference
 * src/decode.c:313-345,388-392: the reader's int32 samples (interleaved, `kind` / `bits` as in blx_resample.h,
 * 1 or 2 channels, in_rate Hz) to int16 / 22 050 Hz / stereo, host to host. out == NULL only returns the
 * number of output frames. */
int blx_resample_to_s16(blx_engine *e, const int32_t *samples, int kind, int bits, int channels, int64_t n_frames,
                        int in_rate, int16_t *out, int64_t out_capacity_frames, int64_t *n_out_frames);
/* The same for samples that are stored as int16 (16-bit PCM files, the CD-audio case): half the host->device bytes and
 * no widening pass on the host. */
int blx_resample_s16_to_s16(blx_engine *e, const int16_t *samples, int channels, int64_t n_frames, int in_rate, int16_t *out,
                            int64_t out_capacity_frames, int64_t *n_out_frames);

/* FLAC frames decoded on the device, one thread per frame (csrc/flacdec.cu; the decode stage of reference
 * src/decode.c:352-427 for FLAC input). `hdr` = the chain of frames the host reader found (n_frames records of 40 bytes,
 * bliss_b200/host/flac_core.h), `first` = the first sample of every frame, `out` = samples * channels interleaved int16
 * (out16 != 0, 16-bit streams) or int32 values on the host. Fails (BLX_ERR_ARG) if any frame does not pass its CRC-16 or
 * does not end where the next begins: the caller then decodes on the host, which resynchronises. */
int blx_flac_decode_frames(blx_engine *e, const uint8_t *file, size_t n_bytes, const void *hdr, const uint64_t *first, int n_frames,
                           int channels, int out16, uint64_t samples, void *out);

/* The same for a 16-bit mono / stereo stream that still needs the decode-stage resampler (CD audio): the decoded PCM stays on
 * the device and goes straight through blx_resample.h; `out` receives int16 / 22 050 Hz / stereo, *n_out_frames frames.
 * out == NULL only returns the number of output frames. */
int blx_flac_decode_resample(blx_engine *e, const uint8_t *file, size_t n_bytes, const void *hdr, const uint64_t *first, int n_frames,
                             int channels, uint64_t samples, int in_rate, int16_t *out, int64_t out_capacity_frames,
                             int64_t *n_out_frames);

/* Diagnostic: how many FLAC streams bl_audio_decode has decoded through the device decoder so far in this process. */
int blx_flac_accelerated_count(void);

/* Envelope intermediates for kernel-level parity tests: hop energies E[m]
 * (reference src/tempo_atk_sort.c:150), 2 * (n_samples / 512) doubles, host. */
int blx_envelope_energy_s16(blx_engine *e, const int16_t *pcm, int n_samples, double *energy);
/* Same through the float32 path (front-end in pass 1, doubled-mono form of the envelope kernel):
 * 2 * ((2 * (n_in / 2)) / 512) doubles. */
int blx_envelope_energy_f32(blx_engine *e, const float *pcm, int64_t n_in, double *energy);

/* Per-bin power spectrum of the frequency analyser BEFORE its scalar epilogue: ps[d] = sum over frames
 * of |X_d|^2, d = 1..255 (reference src/frequency_sort.c:88-93); ps[0] = ps[256] = 0. 257 floats, host. */
int blx_frequency_spectrum_s16(blx_engine *e, const int16_t *pcm, int n_samples, int channels, float *ps);

/* The amplitude analyser's sample histogram as pass 1 counts it: bins for sample values -1904..+1902
 * (3807 counters, value v at index v + 1904), over ALL samples; first/last_nonzero = the bounds of
 * reference src/amplitude_sort.c:26-31 (-1 for an all-zero song). */
int blx_histogram_s16(blx_engine *e, const int16_t *pcm, int n_samples, unsigned *hist /* [3807] */,
                      int *first_nonzero, int *last_nonzero);

/* The envelope analyser's sequential tail ALONE on caller-supplied hop energies (reference
 * src/tempo_atk_sort.c:184-287: log compression, IIR, rectified difference, mix, two box filters, onset
 * count, scores). energy: nb_frames = 2 * (n_samples / 512) doubles (the last two 0), host. */
int blx_envelope_tail(blx_engine *e, const double *energy, int nb_frames, int n_samples, uint64_t duration_s,
                      int *beat, float *tempo, float *attack);

/* ---- measurement ------------------------------------------------------------------
 * With profiling on, every kernel launch is bracketed by CUDA events on its own stream.
 * blx_profile_read synchronises and returns, per kernel id, the accumulated device time
 * (ms) and launch count since the last blx_profile_reset. */
#define BLX_K_PASS1 0    /* front-end + Hann + rFFT-512 + |X|^2 (+ histogram, stats) */
#define BLX_K_EPILOGUE 1 /* band ratios, 301-pass histogram smoothing, mean/variance */
#define BLX_K_ENVELOPE 2 /* normalise + FIR + FP64 rFFT-512 + float-accumulated energy */
#define BLX_K_TAIL 3     /* IIR, diff, box filters, onset count, rating */
#define BLX_K_DISTANCE 4
#define BLX_K_COUNT 5
int blx_profile_enable(blx_engine *e, int on);
int blx_profile_reset(blx_engine *e);
int blx_profile_read(blx_engine *e, float *ms /* [BLX_K_COUNT] */, int *launches /* [BLX_K_COUNT] */);
const char *blx_kernel_name(int kernel_id);

/* The FP64 roofline denominator, measured on this device now: best of several launches of a kernel that
 * only issues independent DFMA chains (2 flop each) from 2 x 1024 threads per SM. MEASURED_PEAKS.json has
 * no FP64 entry, and the envelope kernel is bound by the FP64 pipe. */
int blx_measure_fp64_peak(blx_engine *e, double *tflops, double *sm_mhz_nominal /* may be NULL */);

/* Total kernels launched by this engine since creation (bench.py's gpu_launches). */
long long blx_launch_count(blx_engine *e);

#ifdef __cplusplus
}
#endif
#endif /* BLX_H_ */
