/* C test of the multi-device C-ABI (include/blx.h): one process, every visible B200, no Python.
 *   tests/c/bin/test_multi [max_devices]
 * Songs shard over the devices; the records must be byte-identical to one engine on device 0; the all-gathered
 * nearest-neighbour search (ncclAllGather + distance kernel per device) must agree with a brute-force scan over
 * bl_distance (include/bliss.h, the reference's scalar API) - lowest index on ties. Exit code 0 = pass. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bliss.h"
#include "blx.h"

static unsigned long long rng_state = 0x5EED1234ABCDull;
static double urand(void) {
    rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
    return (double)(rng_state >> 11) / 9007199254740992.0;
}

static int16_t *make_song(int seconds_x10, int *n_out) {
    const int frames = 2205 * seconds_x10, n = 2 * frames;
    int16_t *pcm = (int16_t *)malloc((size_t)n * sizeof(int16_t));
    const double f0 = 80 + 900 * urand(), bpm = 60 + 120 * urand(), amp = 2000 + 9000 * urand(), noise = 300 + 2500 * urand();
    for (int t = 0; t < frames; ++t) {
        const double sec = t / 22050.0, phase = sec * bpm / 60.0, burst = exp(-(phase - floor(phase)) * 8.0);
        const double v = amp * burst * sin(2 * M_PI * f0 * sec) + noise * (urand() + urand() + urand() - 1.5);
        pcm[2 * t] = (int16_t)v;
        pcm[2 * t + 1] = (int16_t)(0.8 * v + 200 * (urand() - 0.5));
    }
    *n_out = n;
    return pcm;
}

#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } } while (0)

static int brute_force_check(const float *v, int n, const int *idx, const float *dist) {
    for (int i = 0; i < n; ++i) {
        int best = -1;
        float bd = 0;
        struct force_vector_s a = {v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]};
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            struct force_vector_s b = {v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]};
            const float d = bl_distance(a, b);
            if (best < 0 || d < bd) { best = j; bd = d; }
        }
        if (idx[i] != best || dist[i] != bd) {
            fprintf(stderr, "row %d: got (%d, %.9g), brute force (%d, %.9g)\n", i, idx[i], dist[i], best, bd);
            return 1;
        }
    }
    return 0;
}

int main(int argc, char **argv) {
    const int max_dev = argc > 1 ? atoi(argv[1]) : 0;
    const int S = 37; /* not a multiple of 2, 4 or 8: ragged blocks */
    int16_t *pcm[37];
    int ns[37], ch[37];
    uint64_t dur[37];
    for (int i = 0; i < S; ++i) {
        const int sx10 = 25 + (int)(urand() * 150);
        pcm[i] = make_song(sx10, &ns[i]);
        ch[i] = 2;
        dur[i] = (uint64_t)(sx10 / 10);
    }
    blx_engine *one = NULL;
    CHECK(blx_init(0, &one) == BLX_OK, "blx_init: %s", blx_last_error());
    blx_result ref[37], got[37];
    CHECK(blx_analyze_batch_s16(one, (const int16_t *const *)pcm, ns, ch, dur, S, BLX_DO_ALL, ref) == BLX_OK, "%s", blx_last_error());
    for (int i = 0; i < S; ++i) CHECK(ref[i].status == 0, "song %d status %d", i, ref[i].status);

    blx_multi *m = NULL;
    CHECK(blx_multi_init(NULL, max_dev, &m) == BLX_OK, "blx_multi_init: %s", blx_last_error());
    const int G = blx_multi_device_count(m);
    printf("devices=%d transport=%s\n", G, blx_multi_transport(m));
    memset(got, 0xff, sizeof(got));
    CHECK(blx_multi_analyze_batch_s16(m, (const int16_t *const *)pcm, ns, ch, dur, S, BLX_DO_ALL, got) == BLX_OK, "%s", blx_last_error());
    CHECK(memcmp(ref, got, sizeof(ref)) == 0, "sharded records differ from the single-device records");
    int taken = 0;
    for (int r = 0; r < G; ++r) { printf("%sdevice %d took %d songs", r ? ", " : "", r, blx_multi_songs_taken(m, r)); taken += blx_multi_songs_taken(m, r); }
    printf("\n");
    CHECK(taken == S, "units taken: %d of %d songs", taken, S);

    /* the chained step: resident vectors -> all-gather -> nearest neighbour per song */
    int idx[37];
    float dist[37], v[37 * 4], g_ms = 0, n_ms = 0;
    CHECK(blx_multi_nearest(m, idx, dist, &g_ms, &n_ms) == BLX_OK, "%s", blx_last_error());
    for (int i = 0; i < S; ++i) { v[4 * i] = ref[i].tempo; v[4 * i + 1] = ref[i].amplitude; v[4 * i + 2] = ref[i].frequency; v[4 * i + 3] = ref[i].attack; }
    CHECK(brute_force_check(v, S, idx, dist) == 0, "nearest neighbours of the analysed batch");
    printf("analysed %d songs on %d device(s): records identical, nearest neighbours = brute force (gather %.3f ms, nearest %.3f ms)\n", S, G, g_ms, n_ms);

    /* a larger table of vectors with exact collisions */
    const int N = 6001;
    float *big = (float *)malloc((size_t)N * 16);
    for (int i = 0; i < 4 * N; ++i) big[i] = (float)((urand() - 0.5) * 40.0);
    memcpy(big + 4 * 4000, big + 4 * 17, 16);   /* song 4000 == song 17 */
    memcpy(big + 4 * 5999, big + 4 * 17, 16);   /* three-way tie: 17's nearest is the lower index, 4000 */
    int *bi = (int *)malloc((size_t)N * 4);
    float *bd = (float *)malloc((size_t)N * 4);
    CHECK(blx_multi_set_vectors(m, big, N) == BLX_OK, "%s", blx_last_error());
    CHECK(blx_multi_nearest(m, bi, bd, &g_ms, &n_ms) == BLX_OK, "%s", blx_last_error());
    CHECK(brute_force_check(big, N, bi, bd) == 0, "nearest neighbours of %d vectors", N);
    CHECK(bi[17] == 4000 && bd[17] == 0.0f && bi[5999] == 17, "tie rule");
    printf("%d vectors on %d device(s): nearest neighbours = brute force over bl_distance (gather %.3f ms, nearest %.3f ms)\n", N, G, g_ms, n_ms);

    blx_multi_shutdown(m);
    blx_shutdown(one);
    for (int i = 0; i < S; ++i) free(pcm[i]);
    free(big); free(bi); free(bd);
    printf("PASS\n");
    return 0;
}
