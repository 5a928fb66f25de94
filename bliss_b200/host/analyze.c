/* bl_analyze, the three analysers and the distance functions of bliss.h as thin wrappers over
 * the device C-ABI (blx.h). Replaces reference src/analyze.c:8-167 and the bodies of
 * src/amplitude_sort.c, src/frequency_sort.c, src/tempo_atk_sort.c. There are no pthreads here:
 * the three analysers are kernels of one stream on the GPU.
 */
#include "../../include/bliss.h"
#include "engine_singleton.h"

static int analyse_song(struct bl_song const *const song, unsigned what, blx_result *r) {
    const int16_t *pcm = (const int16_t *)song->sample_array;
    const int n = song->nSamples;
    const int ch = song->channels;
    const uint64_t dur = song->duration;
    if (!pcm || n <= 0 || (ch != 1 && ch != 2)) {
        fprintf(stderr, "bliss: song has no PCM to analyse\n");
        return BL_UNEXPECTED;
    }
    blx_engine *e = bl_engine_acquire();
    if (!e) return BL_UNEXPECTED;
    int rc = blx_analyze_batch_s16(e, &pcm, &n, &ch, &dur, 1, what, r);
    if (rc != BLX_OK) fprintf(stderr, "bliss: device analysis failed: %s\n", blx_last_error());
    bl_engine_release();
    return rc == BLX_OK ? BL_OK : BL_UNEXPECTED;
}

int bl_analyze(char const *const filename, struct bl_song *current_song) {
    if (bl_audio_decode(filename, current_song) == BL_OK) {
        blx_result r;
        if (analyse_song(current_song, BLX_DO_ALL, &r) != BL_OK) {
            bl_free_song(current_song); /* nothing to rate: do not leave the decoded PCM behind */
            return BL_UNEXPECTED;
        }
        current_song->force_vector.tempo = r.tempo;
        current_song->force_vector.amplitude = r.amplitude;
        current_song->force_vector.frequency = r.frequency;
        current_song->force_vector.attack = r.attack;
        current_song->force = r.force;
        current_song->calm_or_loud = r.calm_or_loud;
        if (r.status != 0) {
            fprintf(stderr, "bliss: song cannot be rated (status 0x%x: too short, silent or constant)\n", r.status);
            bl_free_song(current_song);
            return BL_UNEXPECTED;
        }
        return current_song->calm_or_loud;
    }
    fprintf(stderr, "Couldn't decode song\n");
    return BL_UNEXPECTED;
}

void bl_envelope_sort(struct bl_song const *const song, struct envelope_result_s *result) {
    blx_result r;
    if (analyse_song(song, BLX_DO_ENVELOPE | BLX_DO_AMPLITUDE, &r) != BL_OK) {
        result->tempo = NAN;
        result->attack = NAN;
        return;
    }
    result->tempo = r.tempo;
    result->attack = r.attack;
}

float bl_amplitude_sort(struct bl_song const *const song) {
    blx_result r;
    if (analyse_song(song, BLX_DO_AMPLITUDE, &r) != BL_OK) return NAN;
    return r.amplitude;
}

float bl_frequency_sort(struct bl_song const *const song) {
    blx_result r;
    if (analyse_song(song, BLX_DO_FREQUENCY, &r) != BL_OK) return NAN;
    return r.frequency;
}

/* Scalar, by-value API (7 float operations): evaluated in place with the reference's exact
 * expression shape, reference src/analyze.c:96-100. The batched all-pairs form runs on the GPU
 * (blx_distance_matrix). Built with -ffp-contract=off. */
float bl_distance(struct force_vector_s v_song1, struct force_vector_s v_song2) {
    const struct force_vector_s a = v_song1, b = v_song2;
    float d = (float)sqrt((a.tempo - b.tempo) * (a.tempo - b.tempo) + (a.amplitude - b.amplitude) * (a.amplitude - b.amplitude) +
                          (a.frequency - b.frequency) * (a.frequency - b.frequency) +
                          (a.attack - b.attack) * (a.attack - b.attack));
    return d;
}

float bl_cosine_similarity(struct force_vector_s v_song1, struct force_vector_s v_song2) {
    const struct force_vector_s a = v_song1, b = v_song2;
    float s = (a.tempo * b.tempo + a.amplitude * b.amplitude + a.frequency * b.frequency + a.attack * b.attack) /
              (sqrt(a.tempo * a.tempo + a.amplitude * a.amplitude + a.frequency * a.frequency + a.attack * a.attack) *
               sqrt(b.tempo * b.tempo + b.amplitude * b.amplitude + b.frequency * b.frequency + b.attack * b.attack));
    return s;
}

float bl_distance_file(char const *const filename1, char const *const filename2, struct bl_song *song1,
                       struct bl_song *song2) {
    if (bl_analyze(filename1, song1) != BL_UNEXPECTED && bl_analyze(filename2, song2) != BL_UNEXPECTED)
        return bl_distance(song1->force_vector, song2->force_vector);
    return BL_UNEXPECTED;
}

float bl_cosine_similarity_file(char const *const filename1, char const *const filename2, struct bl_song *song1,
                                struct bl_song *song2) {
    if (bl_analyze(filename1, song1) != BL_UNEXPECTED && bl_analyze(filename2, song2) != BL_UNEXPECTED)
        return bl_cosine_similarity(song1->force_vector, song2->force_vector);
    return BL_UNEXPECTED;
}
