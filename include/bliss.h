/* bliss.h — drop-in public header of the B200-native bliss engine.
 *
 * Declares the same structs, status codes and 15 functions as the reference's
 * include/bliss.h (Polochon-street/bliss @ 20c4536, version 1.2), with identical
 * names, field order, argument order and meaning, so existing callers (C, the cffi
 * binding of reference python/build_bliss.py:35-38, Blissify, leleleplayer) compile
 * and link unchanged against this library's libbliss.so. Struct ABI (LP64):
 * sizeof(struct bl_song) == 120, sizeof(struct force_vector_s) == 16.
 *
 * Like the reference header this one carries no extern "C" block and no code outside declarations, so the
 * text filter of reference python/build_bliss.py:35-38 (drop every line that starts with '#', hand the rest to
 * cffi's cdef) works on it unchanged (tools/build_ref_cffi.py, tests/test_ref_callers.py); C++ callers wrap the
 * include themselves, as they have to with the reference.
 *
 * Differences from the reference header, none of which change the ABI:
 *  - the FFmpeg headers (reference include/bliss.h:5-6) are optional: they are only
 *    needed there for a version shim; the libc headers they used to pull in are
 *    included explicitly;
 *  - the per-song analysis behind bl_analyze / bl_*_sort runs on an NVIDIA B200
 *    through the C-ABI in blx.h; there is no CPU implementation in this library.
 */
#ifndef BL_BLISS_H_
#define BL_BLISS_H_

#include <inttypes.h> /* PRId64: reference examples/analyze.c:40 gets it through the FFmpeg headers */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if defined(__has_include)
#if __has_include(<libavformat/avformat.h>) && defined(BLISS_WITH_LIBAV)
#include <libavformat/avformat.h>
#include <libavutil/md5.h>
#endif
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define BL_VERSION 1.2

/* Return codes (reference include/bliss.h:20-24). BL_OK == BL_LOUD == 0. */
#define BL_LOUD 0
#define BL_CALM 1
#define BL_UNKNOWN 2
#define BL_UNEXPECTED -2
#define BL_OK 0

/* reference include/bliss.h:26-31 — 4 x float, no padding. */
struct force_vector_s {
    float tempo;
    float amplitude;
    float frequency;
    float attack;
};

/* reference include/bliss.h:34-37 */
struct envelope_result_s {
    float tempo;
    float attack;
};

/* reference include/bliss.h:39-47 — thread trampolines of the reference's bl_analyze;
 * kept for source compatibility, unused by this engine. */
struct thread_result_s {
    struct bl_song const *const song;
    float result;
};

struct thread_envelope_result_s {
    struct bl_song const *const song;
    struct envelope_result_s *results;
};

/* reference include/bliss.h:49-67. sample_array holds int16 samples (interleaved
 * L,R) despite its int8_t type; nSamples counts int16 values over all channels. */
struct bl_song {
    float force;
    struct force_vector_s force_vector;
    int8_t *sample_array;
    int channels;
    int nSamples;
    int sample_rate;
    int bitrate;
    int nb_bytes_per_sample;
    int calm_or_loud;
    int resampled;
    uint64_t duration;
    char *filename;
    char *artist;
    char *title;
    char *album;
    char *tracknumber;
    char *genre;
};

/* Decode `filename`, analyse it on the GPU and fill `current_song`.
 * Returns BL_LOUD / BL_CALM / BL_UNKNOWN, or BL_UNEXPECTED if the file could not be
 * decoded or the device analysis failed (reference src/analyze.c:33-86). */
int bl_analyze(char const *const filename, struct bl_song *current_song);

/* bl_analyze both files, then the euclidean distance of their force vectors;
 * (float)BL_UNEXPECTED on failure (reference src/analyze.c:105-125). */
float bl_distance_file(char const *const filename1, char const *const filename2, struct bl_song *song1,
                       struct bl_song *song2);

/* Euclidean distance between two force vectors, float arithmetic, structs by value
 * (reference src/analyze.c:88-103). */
float bl_distance(struct force_vector_s v_song1, struct force_vector_s v_song2);

/* Cosine-similarity counterparts (reference src/analyze.c:127-167). */
float bl_cosine_similarity_file(char const *const filename1, char const *const filename2,
                                struct bl_song *song1, struct bl_song *song2);
float bl_cosine_similarity(struct force_vector_s v_song1, struct force_vector_s v_song2);

/* The three analysers, callable on an already decoded bl_song
 * (reference include/bliss.h:184-217). */
void bl_envelope_sort(struct bl_song const *const song, struct envelope_result_s *result);
float bl_amplitude_sort(struct bl_song const *const song);
float bl_frequency_sort(struct bl_song const *const song);

/* Decode an audio file into int16 / 22 050 Hz / stereo PCM + tags
 * (reference include/bliss.h:234-235). BL_OK or BL_UNEXPECTED. */
int bl_audio_decode(char const *const filename, struct bl_song *const song);

/* Lifecycle helpers (reference include/bliss.h:247-262). */
void bl_free_song(struct bl_song *const song);
float bl_version(void);
void bl_initialize_song(struct bl_song *const song);

/* Integer mean / variance of a sample array and the width-`smooth_width` box filter
 * used by the envelope analyser (reference include/bliss.h:270-290). */
int bl_mean(int16_t *sample_array, int nSamples);
int bl_variance(int16_t *sample_array, int nSamples, int mean);
void bl_rectangular_filter(double *sample_array_out, double *sample_array_in, int nSamples,
                           int smooth_width);

#endif /* BL_BLISS_H_ */
