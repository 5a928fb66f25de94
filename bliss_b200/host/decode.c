/* bl_audio_decode for the drop-in: fills struct bl_song the way
 * reference src/decode.c:27-213,215-349 does — sample_array (C heap, int16 interleaved),
 * nSamples, channels = 2, sample_rate = 22050, nb_bytes_per_sample, bitrate, duration (whole
 * seconds), resampled, filename and the five tag strings with the reference's defaults.
 *
 * Containers: FLAC and RIFF/WAVE via flac_reader.c (the reference uses FFmpeg, absent here).
 * Input that is already int16 / 22 050 Hz is passed through bit for bit, as in the reference
 * (reference src/decode.c:317-318: no resampler is set up). Every other rate / sample format goes
 * through the resampler of include/blx_resample.h on the GPU (blx_resample_to_s16), a restatement of
 * libswresample's default polyphase filter that reproduces the reference's md5 pins of its 48 kHz
 * fixtures (reference tests/test_decode.c:35-36,55-56): stereo stays stereo, mono is up-mixed at -3 dB.
 */
#include "../../include/bliss.h"
#include "engine_singleton.h"
#include "flac_reader.h"
#include "../../include/blx_resample.h"

static char *dup_or(const char *s, const char *fallback) {
    const char *src = s ? s : fallback;
    char *d = (char *)malloc(strlen(src) + 1);
    if (d) strcpy(d, src);
    return d;
}

int blx_flac_decode_frames_emulated(const unsigned char *file, size_t n_bytes, const void *hdr, const unsigned long long *first,
                                    int n_frames, int channels, int out16, unsigned long long samples, void *out);

/* Long FLAC streams are decoded on the device (csrc/flacdec.cu): the reader calls back with the chain of frames it found. */
static int g_flac_accel_ok = 0; /* streams the accelerator decoded (diagnostic; tests assert that it ran) */
int blx_flac_accelerated_count(void) { return __atomic_load_n(&g_flac_accel_ok, __ATOMIC_RELAXED); }

static __thread int t_want_fused = 0; /* bl_audio_decode is the caller: a stream that needs the resampler may get it on the device */
static int flac_on_device_impl(blx_pcm_file *f, const uint8_t *file, size_t n_bytes, const void *hdr, const uint64_t *first, size_t n_frames,
                               int channels, int out16, uint64_t samples, void *out);
static int flac_on_device(blx_pcm_file *f, const uint8_t *file, size_t n_bytes, const void *hdr, const uint64_t *first, size_t n_frames,
                          int channels, int out16, uint64_t samples, void *out) {
    const int rc = flac_on_device_impl(f, file, n_bytes, hdr, first, n_frames, channels, out16, samples, out);
    if (rc == 0) __atomic_add_fetch(&g_flac_accel_ok, 1, __ATOMIC_RELAXED);
    return rc;
}

static int flac_on_device_impl(blx_pcm_file *f, const uint8_t *file, size_t n_bytes, const void *hdr, const uint64_t *first, size_t n_frames,
                               int channels, int out16, uint64_t samples, void *out) {
    if (n_frames > 0x7fffffff) return -1;
    if (getenv("BLX_FLAC_EMULATE")) /* tests without a GPU: the host instance of the device code */
        return blx_flac_decode_frames_emulated(file, n_bytes, hdr, (const unsigned long long *)first, (int)n_frames, channels, out16,
                                               samples, out);
    static int have_gpu = -1; /* asked once: without a device the host threads decode (nothing is printed) */
    if (have_gpu < 0) have_gpu = blx_device_count() > 0;
    if (!have_gpu) return -1;
    blx_engine *e = bl_engine_acquire();
    if (!e) return -1;
    int rc;
    if (t_want_fused && out16 && channels <= 2 && f->sample_rate > 0 && f->sample_rate != BLX_RS_OUT_RATE) {
        /* CD audio and the like: decode and resample in one go, only the 22 050 Hz result comes back */
        int64_t n_out = 0;
        int16_t *pcm = NULL;
        rc = blx_flac_decode_resample(e, file, n_bytes, hdr, first, (int)n_frames, channels, samples, f->sample_rate, NULL, 0, &n_out);
        if (rc == BLX_OK && n_out > 0 && n_out < ((int64_t)1 << 30)) pcm = (int16_t *)malloc((size_t)n_out * 2 * sizeof(int16_t));
        else if (rc == BLX_OK) rc = BLX_ERR_ARG;
        if (pcm) {
            rc = blx_flac_decode_resample(e, file, n_bytes, hdr, first, (int)n_frames, channels, samples, f->sample_rate, pcm, n_out, &n_out);
            if (rc == BLX_OK) { f->resampled16 = pcm; f->resampled_frames = (size_t)n_out; }
            else free(pcm);
        } else if (rc == BLX_OK) {
            rc = BLX_ERR_NOMEM;
        }
    } else {
        rc = blx_flac_decode_frames(e, file, n_bytes, hdr, first, (int)n_frames, channels, out16, samples, out);
    }
    bl_engine_release();
    return rc == BLX_OK ? 0 : -1;
}
__attribute__((constructor)) static void install_flac_accel(void) { blx_flac_accel = flac_on_device; }

int bl_audio_decode(char const *const filename, struct bl_song *const song) {
    blx_pcm_file f;
    bl_initialize_song(song);
    t_want_fused = 1;
    const int read_rc = blx_pcm_file_read(filename, &f);
    t_want_fused = 0;
    if (read_rc != 0) {
        fprintf(stderr, "Couldn't open file: %s (not a readable FLAC or WAV file)\n", filename);
        return BL_UNEXPECTED;
    }
    if (f.sample_rate <= 0 || f.n_frames == 0 || f.n_frames > (size_t)0x3fffffff) { /* nSamples is an int (bliss.h) */
        fprintf(stderr, "Couldn't decode %s: unusable stream parameters (%d Hz, %zu frames)\n", filename, f.sample_rate, f.n_frames);
        blx_pcm_file_free(&f);
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
    /* what FFmpeg's decoder would hand over: FLAC <= 16 bit and 16-bit PCM as int16 (left-justified), 17..32 bit as
     * int32, float PCM as float, 8-bit PCM as unsigned bytes */
    const int is_u8 = f.container == 1 && !f.is_float && f.bits_per_sample == 8;
    const int is_s16 = !f.is_float && !is_u8 && f.bits_per_sample <= 16;
    if (f.resampled16) {
        /* a long FLAC stream that the device decoded AND resampled (flac_on_device) */
        song->sample_array = (int8_t *)f.resampled16;
        song->nSamples = (int)(2 * f.resampled_frames);
        song->resampled = 1;
        f.resampled16 = NULL;
        rc = BL_OK;
    } else if (f.sample_rate == BLX_RS_OUT_RATE && is_s16 && (f.channels == 1 || f.channels == 2)) {
        /* native format: copied through untouched (a mono file keeps its mono sample count while
         * channels reads 2, exactly as reference src/decode.c:187-193 leaves it) */
        const size_t n = f.n_frames * (size_t)f.channels;
        const int shift = 16 - f.bits_per_sample;
        int16_t *pcm = NULL;
        if (f.samples16) { /* 16-bit WAVE or FLAC: the reader's buffer is the song's */
            pcm = f.samples16;
            f.samples16 = NULL;
        } else if ((pcm = (int16_t *)malloc((n ? n : 1) * sizeof(int16_t))) != NULL) {
            for (size_t i = 0; i < n; ++i) pcm[i] = (int16_t)((uint32_t)f.samples[i] << shift);
        }
        if (pcm) {
            song->sample_array = (int8_t *)pcm;
            song->nSamples = (int)n;
            rc = BL_OK;
        }
    } else if (f.channels == 1 || f.channels == 2) {
        /* everything else goes through the resampler (reference src/decode.c:313-345: libswresample to
         * int16 / 22 050 Hz / stereo; here include/blx_resample.h on the GPU) */
        const int kind = f.is_float ? BLX_RS_KIND_F32 : is_u8 ? BLX_RS_KIND_U8 : is_s16 ? BLX_RS_KIND_S16 : BLX_RS_KIND_S32;
        /* 16-bit WAVE and FLAC files arrive as int16 and go to the device as they are; everything else as the reader's int32 */
        const int packed16 = f.samples16 != NULL && !f.samples;
        blx_engine *e = (packed16 || blx_pcm_file_samples32(&f) == 0) ? bl_engine_acquire() : NULL;
        if (e) {
            int64_t n_out = 0;
            int brc = packed16 ? blx_resample_s16_to_s16(e, f.samples16, f.channels, (int64_t)f.n_frames, f.sample_rate, NULL, 0, &n_out)
                               : blx_resample_to_s16(e, f.samples, kind, f.bits_per_sample, f.channels, (int64_t)f.n_frames,
                                                     f.sample_rate, NULL, 0, &n_out);
            int16_t *pcm = NULL;
            if (brc == BLX_OK && n_out > 0 && n_out < ((int64_t)1 << 30)) pcm = (int16_t *)malloc((size_t)n_out * 2 * sizeof(int16_t));
            if (pcm && (packed16 ? blx_resample_s16_to_s16(e, f.samples16, f.channels, (int64_t)f.n_frames, f.sample_rate, pcm, n_out, &n_out)
                                 : blx_resample_to_s16(e, f.samples, kind, f.bits_per_sample, f.channels, (int64_t)f.n_frames,
                                                       f.sample_rate, pcm, n_out, &n_out)) == BLX_OK) {
                song->sample_array = (int8_t *)pcm;
                song->nSamples = (int)(2 * n_out);
                song->resampled = 1;
                rc = BL_OK;
            } else {
                free(pcm);
                if (brc != BLX_OK || n_out > 0) fprintf(stderr, "bliss: resampling failed: %s\n", blx_last_error());
            }
            bl_engine_release();
        }
    } else {
        fprintf(stderr, "Couldn't decode %s: %d channels need a down-mix matrix this build does not carry (mono and stereo only)\n",
                filename, f.channels);
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
