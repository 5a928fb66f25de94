// Host model of float_chain v2 (csrc/envelope.cu): per-lane binade prediction with at most ONE crossing per lane,
// checked against the sequential reference chain s <- (float)((double)s + p_k) on random spectra.
//   gcc -O2 -ffp-contract=off tools/chain_model2.c -lm -o /tmp/chain_model2 && /tmp/chain_model2 600000
// A lane b (1..15) owns bins 16 b + 1 .. 16 b + 16; bins 0..16 run as the plain chain (r16). For lane b:
//   excl = exact double prefix in front of the lane, e0 = its exponent, eE = exponent at the lane's end,
//   x_i  = "the prefix after bin i has left binade e0", decided by comparing the HIGH WORDS of the local prefix
//          c[i] and of thr = 2^(e0+1) - excl (an approximation: everything is verified afterwards),
//   bins with !x_i are converted at grid e0, bins with x_i at grid eE, the first bin with x_i is the lane's crossing
//   and is added by the reference's own double-add / float-convert step.
// Verified: at every crossing q < 2^24 and e0 == the chain's exponent before, eE == its exponent after; lanes
// without a crossing have eE == e0; at the end q < 2^24 and eE(15) == the chain's exponent. Anything else falls
// back to the binade-by-binade scan (not modelled here: counted).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline int hi(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(u >> 32); }
static inline unsigned lo(double d) { uint64_t u; memcpy(&u, &d, 8); return (unsigned)u; }
static inline double mk(int h, unsigned l) { uint64_t u = ((uint64_t)(unsigned)h << 32) | l; double d; memcpy(&d, &u, 8); return d; }
static double ref_chain(const double *p) { float s = 0; for (int k = 0; k <= 256; ++k) s = (float)((double)s + p[k]); return (double)s; }
static long n_cross_total = 0, n_trips_total = 0;
static int fast_chain(const double *p, double *out) {
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
    unsigned T[16], A2[16]; int e0[16], eE[16], ix[16], hasx[16], consistent = 1;
    for (int b = 0; b < 16; ++b) {
        const int h0 = hi(excl[b]), hE = hi(inc[b]);
        e0[b] = (h0 >> 20) & 0x7ff; eE[b] = (hE >> 20) & 0x7ff;
        const int hiM0 = (h0 & 0x7FF00000) + 0x1D80000, hiME = (hE & 0x7FF00000) + 0x1D80000;
        const double thr = mk((h0 & 0x7FF00000) + 0x100000, 0) - excl[b];
        const int thr_hi = (b == 0) ? 0x7fffffff : hi(thr);
        unsigned t = 0, a2 = 0, mask = 0;
        for (int i = 0; i < 16; ++i) {
            const int x = hi(cpre[b][i]) >= thr_hi;
            const double v = p[16 * b + 1 + i] + mk(x ? hiME : hiM0, 0);
            t += lo(v);
            if (x) { a2 += lo(v); mask |= 1u << i; }
        }
        if (b == 0) { t = 0; a2 = 0; mask = 0; }
        T[b] = t; A2[b] = a2; hasx[b] = mask != 0; ix[b] = 16 - __builtin_popcount(mask);
        if (b != 0 && !hasx[b] && eE[b] != e0[b]) consistent = 0;
    }
    unsigned incI[16]; unsigned run = 0;
    for (int b = 0; b < 16; ++b) { run += T[b]; incI[b] = run; }
    int ok = consistent;
    int ex = (hi(r16) >> 20) & 0x7ff;
    if (ex < 1023 - 126 || ex > 1023 + 126) ok = 0;
    unsigned q = ((hi(r16) & 0xFFFFF) << 3) | (lo(r16) >> 29) | 0x800000;
    unsigned Pprev = 0; int trips = 0;
    for (int b = 1; b < 16; ++b) if (hasx[b]) {
        n_cross_total++; trips++;
        const unsigned before = incI[b] - A2[b]; // every bin in front of the crossing bin
        const double pk = p[16 * b + 1 + ix[b]];
        const int hiME = ((eE[b] << 20) & 0x7FF00000) + 0x1D80000;
        const unsigned own = lo(pk + mk(hiME, 0)); // the crossing bin's own increment, counted in A2 / T
        const unsigned qb = q + (before - Pprev);
        if (!(qb < (1u << 24)) || e0[b] != ex) ok = 0;
        const double sq = mk((ex << 20) | ((qb & 0x7FFFFF) >> 3), (qb & 7) << 29);
        const double r = (double)(float)(sq + pk);
        ex = (hi(r) >> 20) & 0x7ff; q = ((hi(r) & 0xFFFFF) << 3) | (lo(r) >> 29) | 0x800000;
        if (ex != eE[b]) ok = 0;
        Pprev = before + own;
    }
    n_trips_total += trips;
    const unsigned qf = q + (incI[15] - Pprev);
    if (!(qf < (1u << 24)) || eE[15] != ex || ex > 1023 + 126) ok = 0;
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
        double f; const int ok = fast_chain(p, &f); const double r = ref_chain(p);
        if (ok) { n_ok++; if (f != r) { n_bad++; if (n_bad < 10) printf("MISMATCH t=%d kind=%d fast=%.17g ref=%.17g\n", t, kind, f, r); } }
        else { n_fb++; fbk[kind]++; }
        nk[kind]++;
    }
    for (int i = 0; i < 8; ++i) printf("kind %d fallback %.5f\n", i, (double)fbk[i] / nk[i]);
    printf("trials %d verified %ld fallback %ld mismatches %ld crossings/hop %.2f\n", N, n_ok, n_fb, n_bad, (double)n_cross_total / N);
    return n_bad != 0;
}
