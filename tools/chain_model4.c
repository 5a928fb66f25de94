// Host model of float_chain (csrc/envelope.cu, "clean / dirty groups" form), checked against the sequential
// reference chain s <- (float)((double)s + p_k), k = 0..256 (reference src/tempo_atk_sort.c:142-149):
//   gcc -O2 -ffp-contract=off -DG=4 tools/chain_model4.c -lm -o /tmp/cm4 && /tmp/cm4 800000
// Lane b (1..15) owns bins 16 b + 1 .. 16 b + 16 in groups of G consecutive bins; bins 0..16 run as the plain chain.
// With S = the exact double prefix sum: a group whose prefix keeps its exponent from the group's start to its end is
// CLEAN - its bins are added as integers on the float grid of that binade (RN_g(p) / g from one double addition with
// a magic constant) - any other group is DIRTY and its G bins are added one by one with the reference's own step.
// Verified on the way: in front of every dirty group the integer sum is below 2^24 and the chain's exponent is the
// predicted one, behind it the chain's exponent is the predicted one again; the same at the end. A hop that fails
// (a prefix sum within rounding noise of a power of two) is redone by the binade-by-binade scan (counted here).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifndef G
#define G 4
#endif
static inline int hi(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(u >> 32); }
static inline unsigned lo(double d) { uint64_t u; memcpy(&u, &d, 8); return (unsigned)u; }
static inline double mk(int h, unsigned l) { uint64_t u = ((uint64_t)(unsigned)h << 32) | l; double d; memcpy(&d, &u, 8); return d; }
static double ref_chain(const double *p) { float s = 0; for (int k = 0; k <= 256; ++k) s = (float)((double)s + p[k]); return (double)s; }
static long n_dirty_total = 0;
static int fast_chain(const double *p, double *out) {
    enum { NG = 16 / G };
    float sf = 0; for (int k = 0; k <= 16; ++k) sf = (float)((double)sf + p[k]);
    const double r16 = (double)sf;
    double tot[16], cpre[16][16], excl[16];
    for (int b = 0; b < 16; ++b) {
        double c = (b == 0) ? p[0] : 0.0;
        for (int i = 0; i < 16; ++i) { c += p[16 * b + 1 + i]; cpre[b][i] = c; }
        tot[b] = c;
    }
    double inc[16]; memcpy(inc, tot, sizeof inc);
    for (int o = 1; o < 16; o <<= 1) { double t[16]; memcpy(t, inc, sizeof t); for (int b = o; b < 16; ++b) inc[b] = t[b] + t[b - o]; }
    for (int b = 0; b < 16; ++b) excl[b] = b ? inc[b - 1] : 0.0;
    if (inc[15] == inc[0]) { *out = r16; return 1; } // nothing left to add after bin 16
    unsigned Gs[16][NG]; int es[16][NG + 1], dirty[16][NG];
    for (int b = 0; b < 16; ++b) {
        int h = hi(excl[b]);
        es[b][0] = (h >> 20) & 0x7ff;
        for (int j = 0; j < NG; ++j) {
            const int hn = hi(j == NG - 1 ? inc[b] : excl[b] + cpre[b][G * j + G - 1]);
            const int hiM = (h & 0x7FF00000) + 0x1D80000;
            unsigned g = 0;
            for (int i = G * j; i < G * j + G; ++i) g += lo(p[16 * b + 1 + i] + mk(hiM, 0));
            dirty[b][j] = (b != 0) && (((hn ^ h) & 0x7FF00000) != 0);
            Gs[b][j] = (dirty[b][j] || b == 0) ? 0u : g;
            es[b][j + 1] = (hn >> 20) & 0x7ff;
            h = hn;
        }
    }
    int ok = 1;
    int ex = (hi(r16) >> 20) & 0x7ff;
    if (ex < 1023 - 126 || ex > 1023 + 126) ok = 0;
    unsigned q = ((hi(r16) & 0xFFFFF) << 3) | (lo(r16) >> 29) | 0x800000;
    unsigned run = 0, Pprev = 0;
    for (int b = 1; b < 16; ++b)
        for (int j = 0; j < NG; ++j) {
            if (!dirty[b][j]) { run += Gs[b][j]; continue; }
            n_dirty_total++;
            const unsigned qb = q + (run - Pprev);
            if (!(qb < (1u << 24)) || es[b][j] != ex) ok = 0;
            double r = mk((ex << 20) | ((qb & 0x7FFFFF) >> 3), (qb & 7) << 29);
            for (int i = G * j; i < G * j + G; ++i) r = (double)(float)(r + p[16 * b + 1 + i]); // reference step
            ex = (hi(r) >> 20) & 0x7ff; q = ((hi(r) & 0xFFFFF) << 3) | (lo(r) >> 29) | 0x800000;
            if (ex != es[b][j + 1] || ex > 1023 + 126) ok = 0;
            Pprev = run;
        }
    const unsigned qf = q + (run - Pprev);
    if (!(qf < (1u << 24)) || es[15][NG] != ex || ex > 1023 + 126) ok = 0;
    *out = mk((ex << 20) | ((qf & 0x7FFFFF) >> 3), (qf & 7) << 29);
    return ok;
}
static double urand(void) { return (rand() + 0.5) / (RAND_MAX + 1.0); }
int main(int argc, char **argv) {
    int N = argc > 1 ? atoi(argv[1]) : 200000; srand(12345);
    long n_ok = 0, n_fb = 0, n_bad = 0; long fbk[8] = {0}, nk[8] = {0};
    for (int t = 0; t < N; ++t) {
        double p[257]; const int kind = t % 8;
        const double scale = pow(10.0, 8 * urand() - 2);
        for (int k = 0; k <= 256; ++k) {
            double x = -log(urand());
            if (kind == 1) x *= 1.0 / (1 + k * 0.05);
            if (kind == 2) x *= (k % 37 == 5) ? 3000.0 : 1.0;
            if (kind == 3) x *= (k > 100) ? 50.0 : 0.01;
            if (kind == 4) x *= exp(-(k / 20.0));
            if (kind == 5) x *= pow(10.0, 6 * urand() - 3);
            if (kind == 6) x *= 1.0 / (1.0 + pow(k / 12.0, 2.0)) * ((k < 128) ? 1.0 : 0.05); // music-like: low-pass + weak image
            if (kind == 7) x *= (k == 20 || k == 27) ? 1e5 : 1.0;                              // two tones inside one lane
            p[k] = x * scale;
        }
        if (t % 97 == 0) for (int k = 0; k <= 256; ++k) p[k] = 0.0; // silence
        if (t % 101 == 0) for (int k = 0; k <= 256; ++k) p[k] *= 1e-42; // subnormal sums
        double f; const int ok = fast_chain(p, &f); const double r = ref_chain(p);
        if (ok) { n_ok++; if (f != r) { n_bad++; if (n_bad < 10) printf("MISMATCH t=%d kind=%d fast=%.17g ref=%.17g\n", t, kind, f, r); } }
        else { n_fb++; fbk[kind]++; }
        nk[kind]++;
    }
    for (int i = 0; i < 8; ++i) printf("kind %d fallback %.5f\n", i, (double)fbk[i] / nk[i]);
    printf("G=%d trials %d verified %ld fallback %ld mismatches %ld dirty groups/hop %.2f\n", G, N, n_ok, n_fb, n_bad, (double)n_dirty_total / N);
    return n_bad != 0;
}
