/* Host-side container readers (FLAC, WAV) used by bl_audio_decode. See flac_reader.c. */
#ifndef BLX_FLAC_READER_H
#define BLX_FLAC_READER_H
#include <stddef.h>
#include <stdint.h>

typedef struct blx_pcm_file {
    int32_t *samples;      /* interleaved, n_frames * channels; raw IEEE bits when is_float; NULL while only samples16 is held */
    int16_t *samples16;    /* 16-bit WAVE and FLAC streams are delivered as int16 (no widening pass); else NULL */
    size_t n_frames;       /* sample frames (per channel) */
    int channels;
    int sample_rate;
    int bits_per_sample;
    int is_float;
    int container;         /* 0 FLAC, 1 RIFF/WAVE */
    uint64_t file_bytes;
    uint8_t md5[16];       /* FLAC STREAMINFO md5 of the unencoded audio (zero for WAV) */
    char *artist, *title, *album, *tracknumber, *genre; /* NULL when absent */
    /* set by the accelerator instead of samples / samples16 when it also ran the decode-stage resampler on the device:
     * int16 / 22 050 Hz / stereo, resampled_frames frames (n_frames, channels, sample_rate still describe the file) */
    int16_t *resampled16;
    size_t resampled_frames;
} blx_pcm_file;

/* Optional accelerator for long FLAC streams: the library (decode.c) points it at the device decoder (one thread per frame,
 * csrc/flacdec.cu); NULL in stand-alone builds of the reader (tools, fuzzer). It gets the file, the chain of frames the
 * scan found (flac_hdr records, flac_core.h) and the first sample of each, and fills `out` (interleaved int16 if out16,
 * else int32) - or f->resampled16, see above; anything but 0 makes the reader decode on the host threads instead. */
typedef int (*blx_flac_accel_fn)(struct blx_pcm_file *f, const uint8_t *file, size_t n_bytes, const void *hdr, const uint64_t *first, size_t n_frames,
                                 int channels, int out16, uint64_t samples, void *out);
extern blx_flac_accel_fn blx_flac_accel;

int blx_pcm_file_read(const char *filename, blx_pcm_file *out); /* 0 on success */
int blx_pcm_file_samples32(blx_pcm_file *f); /* makes f->samples valid (widens samples16 if need be); 0 on success */
void blx_pcm_file_free(blx_pcm_file *f);
#endif
