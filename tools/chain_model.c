// Host model of the predicted-binade float-accumulation chain of csrc/envelope.cu (float_chain), checked against the
// sequential reference chain s <- (float)((double)s + p_k) on random spectra: gcc -O2 -ffp-contract=off tools/chain_model.c -lm
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline int hi(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(u >> 32); }
static inline unsigned lo(double d) { uint64_t u; memcpy(&u, &d, 8); return (unsigned)u; }
static inline double mk(int h, unsigned l) { uint64_t u = ((uint64_t)(unsigned)h << 32) | l; double d; memcpy(&d, &u, 8); return d; }
static double ref_chain(const double *p) { float s = 0; for (int k = 0; k <= 256; ++k) s = (float)((double)s + p[k]); return (double)s; }
static int n_cross_total = 0;
static int fast_chain(const double *p, double *out) {
    float sf = 0; for (int k = 0; k <= 16; ++k) sf = (float)((double)sf + p[k]);
    double r16 = (double)sf;
    double tot[16], excl[16], cpre[16][16];
    for (int b = 0; b < 16; ++b) {
        double c = (b == 0) ? p[0] : 0.0;
        for (int i = 0; i < 16; ++i) { c += p[16 * b + 1 + i]; cpre[b][i] = c; }
        tot[b] = c;
    }
    double inc[16]; memcpy(inc, tot, sizeof inc);
    for (int o = 1; o < 16; o <<= 1) { double t[16]; memcpy(t, inc, sizeof t); for (int b = o; b < 16; ++b) inc[b] = t[b] + t[b - o]; }
    for (int b = 0; b < 16; ++b) excl[b] = b ? inc[b - 1] : 0.0;
    unsigned Pb[258]; int Eg[258]; unsigned char X[258]; memset(X, 0, sizeof X);
    int lanesum[16];
    unsigned Lx[16][17];
    for (int b = 0; b < 16; ++b) {
        int hprev = hi(excl[b]); unsigned acc = 0;
        for (int i = 0; i < 16; ++i) {
            const int k = 16 * b + 1 + i;
            const int hs = hi(excl[b] + cpre[b][i]);
            const int eg = hprev >> 20;
            const int hiM = (hprev & 0x7FF00000) + 0x1D80000;
            const double t = p[k] + mk(hiM, 0);
            const int x = (hs >> 20) != eg;
            unsigned I = x ? 0u : lo(t);
            if (b == 0) { I = 0; }
            Lx[b][i] = acc; acc += I;
            Eg[k] = eg; X[k] = (b == 0) ? 0 : x;
            hprev = hs;
        }
        Lx[b][16] = acc; lanesum[b] = acc;
        if (b == 15) Eg[257] = hprev >> 20;
    }
    unsigned base = 0;
    for (int b = 0; b < 16; ++b) { for (int i = 0; i < 16; ++i) Pb[16 * b + 1 + i] = base + Lx[b][i]; base += lanesum[b]; }
    Pb[257] = base;
    int ok = 1;
    int ex = (hi(r16) >> 20) & 0x7ff;
    if (ex < 1023 - 126 || ex > 1023 + 126) ok = 0;
    unsigned q = ((hi(r16) & 0xFFFFF) << 3) | (lo(r16) >> 29) | 0x800000;
    unsigned Pprev = 0;
    for (int k = 17; k <= 256; ++k) if (X[k]) {
        n_cross_total++;
        const unsigned qb = q + (Pb[k] - Pprev);
        if (!(qb < (1u << 24)) || Eg[k] != ex) ok = 0;
        const double sq = mk((ex << 20) | ((qb & 0x7FFFFF) >> 3), (qb & 7) << 29);
        const double r = (double)(float)(sq + p[k]);
        ex = (hi(r) >> 20) & 0x7ff; q = ((hi(r) & 0xFFFFF) << 3) | (lo(r) >> 29) | 0x800000; Pprev = Pb[k];
    }
    const unsigned qf = q + (Pb[257] - Pprev);
    if (!(qf < (1u << 24)) || Eg[257] != ex || ex > 1023 + 126) ok = 0;
    *out = mk((ex << 20) | ((qf & 0x7FFFFF) >> 3), (qf & 7) << 29);
    return ok;
}
static double urand(void) { return (rand() + 0.5) / (RAND_MAX + 1.0); }
int main(int argc, char **argv) {
    int N = argc > 1 ? atoi(argv[1]) : 200000; srand(12345);
    long n_ok = 0, n_fb = 0, n_bad = 0; long fbk[6] = {0}, nk[6] = {0};
    for (int t = 0; t < N; ++t) {
        double p[257]; const int kind = t % 6;
        const double scale = pow(10.0, 8 * urand() - 2);
        for (int k = 0; k <= 256; ++k) {
            double x = -log(urand());
            if (kind == 1) x *= 1.0 / (1 + k * 0.05);
            if (kind == 2) x *= (k % 37 == 5) ? 3000.0 : 1.0;
            if (kind == 3) x *= (k > 100) ? 50.0 : 0.01;
            if (kind == 4) x *= exp(-(k / 20.0));
            if (kind == 5) x *= pow(10.0, 6 * urand() - 3);
            p[k] = x * scale;
        }
        double f; const int ok = fast_chain(p, &f); const double r = ref_chain(p);
        if (ok) { n_ok++; if (f != r) { n_bad++; if (n_bad < 10) printf("MISMATCH t=%d kind=%d fast=%.17g ref=%.17g\n", t, kind, f, r); } }
        else { n_fb++; fbk[kind]++; }
        nk[kind]++;
    }
    for (int i = 0; i < 6; ++i) printf("kind %d fallback %.5f\n", i, (double)fbk[i] / nk[i]);
    printf("trials %d verified %ld fallback %ld mismatches %ld crossings/hop %.2f\n", N, n_ok, n_fb, n_bad, (double)n_cross_total / N);
    return n_bad != 0;
}
