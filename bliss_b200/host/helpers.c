/* Lifecycle helpers and small numeric helpers of bliss.h.
 *
 * bl_free_song / bl_initialize_song / bl_version mirror reference src/helpers.c:3-28 (same
 * ownership contract: the seven pointers come from the C heap and are NULLed after free).
 * bl_mean / bl_variance (reference src/helpers.c:30-49) and bl_rectangular_filter (reference
 * src/tempo_atk_sort.c:19-40) run on the GPU through blx.h like the analysers that use them.
 */
#include <pthread.h>

#include "../../include/bliss.h"
#include "engine_singleton.h"

/* A small pool of engines (one per concurrent caller, each with its own streams and device buffers): bl_analyze
 * and the stand-alone analysers may be called from several threads at once - unlike the reference, whose fftw
 * planner calls make that unsafe (reference src/tempo_atk_sort.c:94,295) - and their kernels then overlap on the
 * GPU. The mutex only guards the pool's free list; no caller holds it while it analyses. */
#define BL_POOL_MAX 16
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_freed = PTHREAD_COND_INITIALIZER;
static blx_engine *g_pool[BL_POOL_MAX];
static int g_busy[BL_POOL_MAX];
static int g_created = 0, g_failed = 0;
static __thread int t_slot = -1; /* the engine this thread holds between acquire and release */

blx_engine *bl_engine_acquire(void) {
    pthread_mutex_lock(&g_lock);
    for (;;) {
        if (g_failed) break;
        int free_slot = -1;
        for (int i = 0; i < g_created; ++i)
            if (!g_busy[i]) { free_slot = i; break; }
        if (free_slot < 0 && g_created < BL_POOL_MAX) {
            const char *dev = getenv("BLISS_DEVICE");
            blx_engine *e = NULL;
            /* created under the lock: engine start-up is rare and blx_init sets per-device kernel attributes */
            if (blx_init(dev ? atoi(dev) : 0, &e) != BLX_OK) {
                if (g_created == 0) {
                    fprintf(stderr, "bliss: cannot start the GPU engine: %s\n", blx_last_error());
                    g_failed = 1;
                    break;
                }
            } else {
                g_pool[g_created] = e;
                g_busy[g_created] = 0;
                free_slot = g_created++;
            }
        }
        if (free_slot >= 0) {
            g_busy[free_slot] = 1;
            t_slot = free_slot;
            pthread_mutex_unlock(&g_lock);
            return g_pool[free_slot];
        }
        pthread_cond_wait(&g_freed, &g_lock); /* BL_POOL_MAX callers are inside already */
    }
    pthread_mutex_unlock(&g_lock);
    return NULL;
}

void bl_engine_release(void) {
    pthread_mutex_lock(&g_lock);
    if (t_slot >= 0) g_busy[t_slot] = 0;
    t_slot = -1;
    pthread_cond_signal(&g_freed);
    pthread_mutex_unlock(&g_lock);
}

void bl_free_song(struct bl_song *const song) {
    free(song->artist);
    free(song->title);
    free(song->album);
    free(song->tracknumber);
    free(song->genre);
    free(song->filename);
    free(song->sample_array);
    bl_initialize_song(song);
}

void bl_initialize_song(struct bl_song *const song) {
    song->sample_array = NULL;
    song->filename = NULL;
    song->artist = NULL;
    song->title = NULL;
    song->album = NULL;
    song->tracknumber = NULL;
    song->genre = NULL;
}

float bl_version(void) {
    printf("Using bliss analyzer version %0.1f.\n", BL_VERSION);
    return (float)BL_VERSION;
}

int bl_mean(int16_t *sample_array, int nSamples) {
    int mean = 0;
    blx_engine *e = bl_engine_acquire();
    if (!e) return 0;
    if (blx_mean_variance_s16(e, sample_array, nSamples, NULL, &mean, NULL) != BLX_OK)
        fprintf(stderr, "bliss: bl_mean failed: %s\n", blx_last_error());
    bl_engine_release();
    return mean;
}

int bl_variance(int16_t *sample_array, int nSamples, int mean) {
    int var = 0;
    blx_engine *e = bl_engine_acquire();
    if (!e) return 0;
    if (blx_mean_variance_s16(e, sample_array, nSamples, &mean, NULL, &var) != BLX_OK)
        fprintf(stderr, "bliss: bl_variance failed: %s\n", blx_last_error());
    bl_engine_release();
    return var;
}

void bl_rectangular_filter(double *sample_array_out, double *sample_array_in, int nSamples, int smooth_width) {
    blx_engine *e = bl_engine_acquire();
    if (!e) return;
    if (blx_rectangular_filter(e, sample_array_out, sample_array_in, nSamples, smooth_width) != BLX_OK)
        fprintf(stderr, "bliss: bl_rectangular_filter failed: %s\n", blx_last_error());
    bl_engine_release();
}
