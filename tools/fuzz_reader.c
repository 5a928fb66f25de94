/* fuzz_reader.c — mutation fuzzing of the host file reader (bliss_b200/host/flac_reader.c), which parses untrusted
 * files. Build with AddressSanitizer + UBSan and run over the FLAC / WAVE fixtures:
 *   gcc -std=gnu99 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=all -Ibliss_b200/host \
 *       -o /tmp/fuzz_reader tools/fuzz_reader.c bliss_b200/host/flac_reader.c -lm
 *   /tmp/fuzz_reader 20000 tests/golden/song.flac tests/golden/song_s32_mono.flac some.wav
 * Every iteration copies a seed file, applies 1..8 mutations (byte flips, 0x00 / 0xFF runs, truncation, a duplicated
 * slice, a 32-bit field set to an extreme value), writes it to a memory-backed temp file and decodes it. Any crash,
 * out-of-bounds access or undefined operation aborts the run; rejected files are the expected outcome. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "flac_reader.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void) {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 16);
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: fuzz_reader iterations file...\n"); return 2; }
    const long iters = atol(argv[1]);
    const int n_seeds = argc - 2;
    uint8_t **seed = calloc(n_seeds, sizeof(*seed));
    size_t *len = calloc(n_seeds, sizeof(*len));
    for (int i = 0; i < n_seeds; ++i) {
        FILE *f = fopen(argv[2 + i], "rb");
        if (!f) { perror(argv[2 + i]); return 2; }
        fseek(f, 0, SEEK_END); len[i] = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
        if (len[i] > (1u << 19)) len[i] = 1u << 19; /* the head of the file is where the structure is; keeps a run short */
        seed[i] = malloc(len[i]);
        if (fread(seed[i], 1, len[i], f) != len[i]) return 2;
        fclose(f);
    }
    char path[64];
    snprintf(path, sizeof(path), "%s/blx_fuzz_%d.bin", access("/dev/shm", W_OK) == 0 ? "/dev/shm" : "/tmp", (int)getpid());
    long accepted = 0, rejected = 0;
    for (long it = 0; it < iters; ++it) {
        const int s = (int)(rnd() % (uint32_t)n_seeds);
        size_t n = len[s];
        uint8_t *b = malloc(n + 4096);
        memcpy(b, seed[s], n);
        const int muts = 1 + (int)(rnd() % 8);
        for (int m = 0; m < muts && n > 16; ++m) {
            /* most mutations land in the first 8 KB (headers, first frames) */
            const size_t pos = (rnd() % 4) ? rnd() % (n < 8192 ? n : 8192) : rnd() % n;
            switch (rnd() % 6) {
            case 0: b[pos] ^= (uint8_t)(1u << (rnd() % 8)); break;
            case 1: b[pos] = (uint8_t)rnd(); break;
            case 2: { size_t run = 1 + rnd() % 64; if (pos + run > n) run = n - pos; memset(b + pos, (rnd() & 1) ? 0xFF : 0x00, run); break; }
            case 3: n = pos > 16 ? pos : n; break; /* truncate */
            case 4: { size_t run = 1 + rnd() % 256; if (pos + 2 * run <= n) memmove(b + pos + run, b + pos, run); break; }
            default: if (pos + 4 <= n) { const uint32_t v = (rnd() & 1) ? 0xFFFFFFFFu : (rnd() & 1) ? 0x7FFFFFFFu : 0u; memcpy(b + pos, &v, 4); } break;
            }
        }
        FILE *f = fopen(path, "wb");
        if (!f) { perror(path); return 2; }
        fwrite(b, 1, n, f);
        fclose(f);
        free(b);
        blx_pcm_file pf;
        if (blx_pcm_file_read(path, &pf) == 0) {
            /* touch everything a caller would */
            volatile long long acc = 0;
            if (blx_pcm_file_samples32(&pf) == 0)
                for (size_t i = 0; i < pf.n_frames * (size_t)pf.channels; i += 97) acc += pf.samples[i];
            if (pf.title) acc += (long long)strlen(pf.title);
            (void)acc;
            blx_pcm_file_free(&pf);
            ++accepted;
        } else {
            ++rejected;
        }
    }
    unlink(path);
    for (int i = 0; i < n_seeds; ++i) free(seed[i]);
    free(seed);
    free(len);
    printf("%ld mutated files: %ld decoded, %ld rejected, no memory or undefined-behaviour error\n", iters, accepted, rejected);
    return 0;
}
