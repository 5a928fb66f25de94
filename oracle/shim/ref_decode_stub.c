/* TEST INFRASTRUCTURE ONLY (oracle/): link-time replacement for the reference's
 * bl_audio_decode (src/decode.c:27-213), which needs FFmpeg. The reference's
 * bl_analyze (src/analyze.c:39) calls it; here it hands over PCM that the test
 * harness registered beforehand, in the exact shape decode.c leaves behind
 * (src/decode.c:187-193: int16, 22 050 Hz, 2 channels, nSamples over both
 * channels, duration in whole seconds).
 */
#include <stddef.h>
#include "bliss.h"

static const int16_t *g_pcm;
static int g_n;
static uint64_t g_duration;
static int g_channels = 2;

void oracle_ref_set_pcm(const int16_t *pcm, int n_samples, uint64_t duration_s, int channels) {
    g_pcm = pcm;
    g_n = n_samples;
    g_duration = duration_s;
    g_channels = channels;
}

int bl_audio_decode(char const *const filename, struct bl_song *const song) {
    (void)filename;
    bl_initialize_song(song);
    if (!g_pcm || g_n <= 0) return BL_UNEXPECTED;
    song->sample_array = malloc((size_t)g_n * sizeof(int16_t));
    if (!song->sample_array) return BL_UNEXPECTED;
    memcpy(song->sample_array, g_pcm, (size_t)g_n * sizeof(int16_t));
    song->nSamples = g_n;
    song->channels = g_channels;
    song->sample_rate = 22050;
    song->nb_bytes_per_sample = 2;
    song->bitrate = 0;
    song->resampled = 0;
    song->duration = g_duration;
    return BL_OK;
}

/* struct layout probes for tests. */
size_t oracle_ref_sizeof_bl_song(void) { return sizeof(struct bl_song); }
size_t oracle_ref_offsetof_sample_array(void) { return offsetof(struct bl_song, sample_array); }
size_t oracle_ref_offsetof_duration(void) { return offsetof(struct bl_song, duration); }
