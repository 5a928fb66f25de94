/* bl_audio_decode for the drop-in: fills struct bl_song the way
 * reference src/decode.c:27-213,215-349 does — sample_array (C heap, int16 interleaved),
 * nSamples, channels = 2, sample_rate = 22050, nb_bytes_per_sample, bitrate, duration (whole
 * seconds), resampled, filename and the five tag strings with the reference's defaults.
 *
 * Containers: FLAC and RIFF/WAVE via flac_reader.c (the reference uses FFmpeg, absent here).
 * Input that is already int16 / 22 050 Hz is passed through bit for bit, as in the reference
 * (reference src/decode.c:317-318: no resampler is set up). Other rates / sample formats need a
 * resampler; FFmpeg's libswresample is third-party and cannot be matched bit for bit, so they are
 * converted by the engine's own front-end where it applies (44.1 kHz, any bit depth -> mono mix ->
 * blx_frontend.h on the GPU) and rejected otherwise.
 */
#include "../../include/bliss.h"
#include "engine_singleton.h"
#include "flac_reader.h"

static char *dup_or(const char *s, const char *fallback) {
    const char *src = s ? s : fallback;
    char *d = (char *)malloc(strlen(src) + 1);
    if (d) strcpy(d, src);
    return d;
}

int bl_audio_decode(char const *const filename, struct bl_song *const song) {
    blx_pcm_file f;
    bl_initialize_song(song);
    if (blx_pcm_file_read(filename, &f) != 0) {
        fprintf(stderr, "Couldn't open file: %s (not a readable FLAC or WAV file)\n", filename);
        return BL_UNEXPECTED;
    }
    const double seconds = (double)f.n_frames / (double)f.sample_rate;
    song->filename = dup_or(filename, "");
    song->duration = (uint64_t)seconds;                       /* reference src/decode.c:235 */
    song->bitrate = seconds > 0 ? (int)((double)f.file_bytes * 8.0 / seconds) : 0; /* libavformat's estimate */
    song->tracknumber = dup_or(f.tracknumber, "");            /* reference src/decode.c:261-270 */
    if (song->tracknumber) song->tracknumber[strcspn(song->tracknumber, "/")] = '\0';
    song->title = dup_or(f.title, "<no title>");
    song->artist = dup_or(f.artist, "<no artist>");
    song->album = dup_or(f.album, "<no album>");
    song->genre = dup_or(f.genre, "<no genre>");
    song->nb_bytes_per_sample = 2;
    song->sample_rate = 22050;                                /* reference src/decode.c:191-193 */
    song->channels = 2;
    song->resampled = 0;

    int rc = BL_UNEXPECTED;
    if (f.sample_rate == 22050 && f.bits_per_sample == 16 && !f.is_float && (f.channels == 1 || f.channels == 2)) {
        /* native format: copied through untouched (a mono file keeps its mono sample count while
         * channels reads 2, exactly as reference src/decode.c:187-193 leaves it) */
        const size_t n = f.n_frames * (size_t)f.channels;
        int16_t *pcm = (int16_t *)malloc(n * sizeof(int16_t));
        if (pcm) {
            for (size_t i = 0; i < n; ++i) pcm[i] = (int16_t)f.samples[i];
            song->sample_array = (int8_t *)pcm;
            song->nSamples = (int)n;
            rc = BL_OK;
        }
    } else if (f.sample_rate == 44100 && f.channels >= 1) {
        /* 44.1 kHz: mono mix on the host (I/O stage), 2:1 front-end on the GPU */
        const size_t n = f.n_frames;
        float *mono = (float *)malloc(n * sizeof(float));
        int16_t *pcm = (int16_t *)malloc((n / 2) * 2 * sizeof(int16_t) + 4);
        if (mono && pcm && n >= 2) {
            const float scale = f.is_float ? 1.0f : 1.0f / (float)(1u << (f.bits_per_sample - 1));
            for (size_t i = 0; i < n; ++i) {
                float acc = 0.0f;
                for (int c = 0; c < f.channels; ++c) {
                    const int32_t raw = f.samples[i * (size_t)f.channels + (size_t)c];
                    float v;
                    if (f.is_float) memcpy(&v, &raw, 4);
                    else v = (float)raw * scale;
                    acc += v;
                }
                mono[i] = acc / (float)f.channels;
            }
            blx_engine *e = bl_engine_acquire();
            if (e) {
                if (blx_frontend_f32(e, mono, (int64_t)n, pcm) == BLX_OK) {
                    song->sample_array = (int8_t *)pcm;
                    song->nSamples = (int)((n / 2) * 2);
                    song->resampled = 1;
                    pcm = NULL;
                    rc = BL_OK;
                } else {
                    fprintf(stderr, "bliss: front-end failed: %s\n", blx_last_error());
                }
                bl_engine_release();
            }
        }
        free(mono);
        free(pcm);
    } else {
        fprintf(stderr,
                "Couldn't decode %s: %d Hz / %d bit needs a resampler this build does not carry "
                "(supported: 22050 Hz s16 passthrough, 44100 Hz via the GPU front-end)\n",
                filename, f.sample_rate, f.bits_per_sample);
    }
    blx_pcm_file_free(&f);
    if (rc != BL_OK) {
        bl_free_song(song);
        return BL_UNEXPECTED;
    }
    if (song->nSamples <= 0) {
        fprintf(stderr, "Couldn't find any valid samples while decoding\n");
        bl_free_song(song);
        return BL_UNEXPECTED;
    }
    return BL_OK;
}
