/* Host-side FLAC / WAV reader for bl_audio_decode.
 *
 * The reference demuxes/decodes with FFmpeg (reference src/decode.c:27-213); this
 * container has no FFmpeg, so the drop-in carries its own small decoders for the
 * container formats its fixtures and benchmarks use: native FLAC (all subframe
 * types, 8..24 bit, 1..2 channels used here, up to 8 decoded) and RIFF/WAVE PCM
 * (s16 / s24 / s32 / f32). Output: interleaved int32 samples at the file's own
 * bit depth + STREAMINFO + Vorbis comments. Conversion to the analysers' input
 * format (int16, 22 050 Hz, stereo) happens in decode.c.
 *
 * Written from the FLAC format specification (RFC 9639); not derived from
 * libFLAC or FFmpeg source.
 */
#include "flac_reader.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

/* ------------------------------------------------------------------ */
/* FLAC                                                                */
/* ------------------------------------------------------------------ */
/* Frame checksums (RFC 9639 section 9.1.8 / 9.3): CRC-8 of the header (polynomial x^8 + x^2 + x + 1) and CRC-16 of
 * the whole frame (x^16 + x^15 + x^2 + 1), both MSB first with a zero start value. A sync code inside damaged or
 * foreign data is only accepted as a frame when both match. */
static uint8_t crc8_of(const uint8_t *p, size_t n) {
    uint8_t c = 0;
    for (size_t i = 0; i < n; ++i) {
        c ^= p[i];
        for (int k = 0; k < 8; ++k) c = (uint8_t)((c & 0x80) ? (c << 1) ^ 0x07 : (c << 1));
    }
    return c;
}
/* CRC-16, eight bytes per step (slicing-by-8): tab[k][b] = the CRC of byte b followed by k zero bytes. The tables are
 * filled once when the library is loaded (no lazy initialisation that concurrent callers could race on). */
static uint16_t crc16_tab[8][256];
__attribute__((constructor)) static void crc16_init(void) {
    for (int i = 0; i < 256; ++i) {
        uint16_t c = (uint16_t)(i << 8);
        for (int k = 0; k < 8; ++k) c = (uint16_t)((c & 0x8000) ? (c << 1) ^ 0x8005 : (c << 1));
        crc16_tab[0][i] = c;
    }
    for (int k = 1; k < 8; ++k)
        for (int i = 0; i < 256; ++i) {
            const uint16_t c = crc16_tab[k - 1][i];
            crc16_tab[k][i] = (uint16_t)((c << 8) ^ crc16_tab[0][c >> 8]);
        }
}
static uint16_t crc16_of(const uint8_t *p, size_t n) {
    uint16_t c = 0;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        const uint8_t b0 = (uint8_t)(p[i] ^ (c >> 8)), b1 = (uint8_t)(p[i + 1] ^ (c & 0xFF));
        c = (uint16_t)(crc16_tab[7][b0] ^ crc16_tab[6][b1] ^ crc16_tab[5][p[i + 2]] ^ crc16_tab[4][p[i + 3]] ^ crc16_tab[3][p[i + 4]] ^
                       crc16_tab[2][p[i + 5]] ^ crc16_tab[1][p[i + 6]] ^ crc16_tab[0][p[i + 7]]);
    }
    for (; i < n; ++i) c = (uint16_t)((c << 8) ^ crc16_tab[0][(c >> 8) ^ p[i]]);
    return c;
}

static void lpc_restore(int32_t *out, int n, int order, const int32_t *coef, int shift, int narrow);
#define FLAC_LPC lpc_restore
#define FLAC_CRC16 crc16_of
#include "flac_core.h"

/* LPC synthesis out[i] += (sum_j coef[j] out[i - 1 - j]) >> shift, i = order..n-1 (RFC 9639 section 9.2.6), the inner
 * loop of the decoder: the taps are reversed once so that the dot product runs over contiguous samples, the common
 * orders get a loop with a constant trip count (unrolled and vectorised by the compiler; an AVX2 clone is picked at
 * load time where the CPU has it), and 32-bit accumulators are used when bps + precision + log2(order) <= 32 bits
 * guarantees that nothing overflows them. */

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__) && !defined(__SANITIZE_THREAD__) /* (TSan cannot run ifunc resolvers) */
#define BLX_CLONES __attribute__((target_clones("avx2", "default"), optimize("O3")))
#else
#define BLX_CLONES
#endif
BLX_CLONES static void lpc_restore(int32_t *out, int n, int order, const int32_t *coef, int shift, int narrow) {
    fc_lpc_unrolled(out, n, order, coef, shift, narrow); /* inlined into each clone */
}

static void add_tag(blx_pcm_file *f, const char *kv, size_t len) {
    const char *eq = memchr(kv, '=', len);
    if (!eq) return;
    size_t klen = (size_t)(eq - kv), vlen = len - klen - 1;
    static const char *keys[5] = {"TRACKNUMBER", "TITLE", "ARTIST", "ALBUM", "GENRE"};
    char **dst[5] = {&f->tracknumber, &f->title, &f->artist, &f->album, &f->genre};
    for (int i = 0; i < 5; ++i) {
        if (strlen(keys[i]) == klen && strncasecmp(kv, keys[i], klen) == 0 && !*dst[i]) {
            char *v = (char *)malloc(vlen + 1);
            if (!v) return;
            memcpy(v, eq + 1, vlen);
            v[vlen] = '\0';
            *dst[i] = v;
        }
    }
}

/* ---- frames ---------------------------------------------------------------------------------------------------- */
/* Parses and checks (CRC-8, channel count) the frame header at d[pos]; 0 if there is a plausible one. */
static int parse_frame_header(const uint8_t *d, size_t n, size_t pos, const blx_pcm_file *f, flac_hdr *h) {
    if (pos + 6 > n || d[pos] != 0xFF || (d[pos + 1] & 0xFE) != 0xF8) return -1;
    bitrd b;
    br_init(&b, d, n, pos);
    br_u(&b, 15);
    h->variable = (int)br_u(&b, 1);
    const int bs_code = (int)br_u(&b, 4);
    const int sr_code = (int)br_u(&b, 4);
    h->ch_code = (int)br_u(&b, 4);
    const int ss_code = (int)br_u(&b, 3);
    br_u(&b, 1);
    /* UTF-8 style coded frame / sample number (up to 36 bits) */
    const uint32_t first = br_u(&b, 8);
    int extra = 0;
    uint64_t num = first;
    if (first & 0x80) {
        uint32_t m = 0x40;
        while (first & m) { extra++; m >>= 1; }
        if (extra == 0 || extra > 6) return -1;
        num = first & (m - 1);
    }
    for (int i = 0; i < extra; ++i) num = (num << 6) | (br_u(&b, 8) & 0x3F);
    h->number = num;
    if (bs_code == 1) h->blocksize = 192;
    else if (bs_code >= 2 && bs_code <= 5) h->blocksize = 576 << (bs_code - 2);
    else if (bs_code == 6) h->blocksize = (int)br_u(&b, 8) + 1;
    else if (bs_code == 7) h->blocksize = (int)br_u(&b, 16) + 1;
    else if (bs_code >= 8) h->blocksize = 256 << (bs_code - 8);
    else return -1;
    if (sr_code == 12) br_u(&b, 8);
    else if (sr_code == 13 || sr_code == 14) br_u(&b, 16);
    const size_t hdr_end = br_bytepos(&b);
    const uint32_t crc8 = br_u(&b, 8);
    if (br_err(&b) || hdr_end > n || crc8 != crc8_of(d + pos, hdr_end - pos)) return -1; /* a false sync */
    static const int ss_tab[8] = {0, 8, 12, 0, 16, 20, 24, 32};
    h->bps = ss_tab[ss_code] ? ss_tab[ss_code] : f->bits_per_sample;
    const int nch = (h->ch_code < 8) ? h->ch_code + 1 : 2;
    if (h->ch_code > 10 || nch != f->channels) return -1;
    h->off = pos;
    h->hdr_len = hdr_end + 1 - pos;
    return 0;
}

/* 16-bit streams are delivered as int16 (blx_pcm_file.samples16): half the memory traffic on the host and on the way
 * to the device */
static void interleave16(int16_t *dst, const int32_t *chbuf, int blocksize, int nch) {
    if (nch == 2) {
        const int32_t *c0 = chbuf, *c1 = chbuf + 65536;
        for (int i = 0; i < blocksize; ++i) { dst[2 * i] = (int16_t)c0[i]; dst[2 * i + 1] = (int16_t)c1[i]; }
    } else {
        for (int i = 0; i < blocksize; ++i)
            for (int c = 0; c < nch; ++c) dst[(size_t)i * (size_t)nch + (size_t)c] = (int16_t)chbuf[(size_t)c * 65536 + (size_t)i];
    }
}
static void interleave(int32_t *dst, const int32_t *chbuf, int blocksize, int nch) {
    if (nch == 2) {
        const int32_t *c0 = chbuf, *c1 = chbuf + 65536;
        for (int i = 0; i < blocksize; ++i) { dst[2 * i] = c0[i]; dst[2 * i + 1] = c1[i]; }
    } else {
        for (int i = 0; i < blocksize; ++i)
            for (int c = 0; c < nch; ++c) dst[(size_t)i * (size_t)nch + (size_t)c] = chbuf[(size_t)c * 65536 + (size_t)i];
    }
}

blx_flac_accel_fn blx_flac_accel = NULL;

/* ---- parallel decode: the frames of a stream are independent ------------------------------------------------------ */
typedef struct {
    const uint8_t *d;
    size_t n;
    const flac_hdr *hdr;     /* the chain of frames found by the scan */
    const size_t *first;     /* first sample (per channel) of every frame */
    size_t n_frames;
    int nch;
    void *pcm;               /* int32, or int16 for a 16-bit stream */
    int out16;
    size_t next;             /* next frame nobody has taken (guarded by lock) */
    int failed;              /* a frame did not decode or did not end where the next one starts */
    pthread_mutex_t lock;
} flac_job;

static void *flac_worker(void *arg) {
    flac_job *j = (flac_job *)arg;
    int32_t *chbuf = (int32_t *)malloc((size_t)65536 * 8 * sizeof(int32_t));
    for (;;) {
        pthread_mutex_lock(&j->lock);
        const size_t lo = j->next;
        size_t hi = lo + 16; /* a batch of frames per visit */
        if (hi > j->n_frames) hi = j->n_frames;
        j->next = hi;
        const int stop = j->failed || !chbuf;
        if (!chbuf) j->failed = 1;
        pthread_mutex_unlock(&j->lock);
        if (lo >= hi || stop) break;
        for (size_t k = lo; k < hi; ++k) {
            size_t end = 0;
            const size_t want = (k + 1 < j->n_frames) ? j->hdr[k + 1].off : 0;
            if (decode_frame(j->d, j->n, &j->hdr[k], chbuf, 65536, &end) != 0 || (want && end != want)) {
                pthread_mutex_lock(&j->lock);
                j->failed = 1;
                pthread_mutex_unlock(&j->lock);
                break;
            }
            if (j->out16) interleave16((int16_t *)j->pcm + j->first[k] * (size_t)j->nch, chbuf, j->hdr[k].blocksize, j->nch);
            else interleave((int32_t *)j->pcm + j->first[k] * (size_t)j->nch, chbuf, j->hdr[k].blocksize, j->nch);
        }
    }
    free(chbuf);
    return NULL;
}

static int flac_decode_threads(void) {
    const char *e = getenv("BLX_DECODE_THREADS");
    int t = e ? atoi(e) : 4;
    return t < 1 ? 1 : (t > 32 ? 32 : t);
}

/* Finds the chain of frames by their headers alone (sync code, CRC-8, coded numbers that continue the sequence), decodes
 * them on several threads and checks on the way that every frame is good and ends exactly where the next begins.
 * 0 = done (f filled); anything unusual (a damaged or foreign byte range, a false sync that fits the sequence, numbers that
 * jump) returns 1 and the caller decodes the stream sequentially, which resynchronises frame by frame. */
static int decode_flac_parallel(const uint8_t *d, size_t n, size_t pos, uint64_t total, blx_pcm_file *f, int threads) {
    size_t cap = 1024, nf = 0;
    flac_hdr *hdr = (flac_hdr *)malloc(cap * sizeof(flac_hdr));
    size_t *first = NULL;
    int rc = 1;
    if (!hdr) return 1;
    size_t samples = 0;
    while (pos + 6 <= n) {
        const uint8_t *q = (const uint8_t *)memchr(d + pos, 0xFF, n - pos);
        if (!q) break;
        pos = (size_t)(q - d);
        flac_hdr h;
        if (parse_frame_header(d, n, pos, f, &h) != 0) { pos++; continue; }
        if (nf) { /* must continue the sequence: a sync code inside a frame's data almost never does */
            const flac_hdr *p = &hdr[nf - 1];
            const uint64_t expect = p->variable ? p->number + (uint64_t)p->blocksize : p->number + 1;
            if (h.variable != p->variable || h.number != expect) { pos++; continue; }
        }
        if (nf == cap) {
            cap *= 2;
            flac_hdr *nh = (flac_hdr *)realloc(hdr, cap * sizeof(flac_hdr));
            if (!nh) goto out;
            hdr = nh;
        }
        hdr[nf++] = h;
        samples += (size_t)h.blocksize;
        pos += h.hdr_len; /* (a frame is longer than its header: the next candidate cannot start inside it) */
    }
    if (nf < 32 || samples > ((size_t)1 << 40)) goto out;
    first = (size_t *)malloc(nf * sizeof(size_t));
    if (!first) goto out;
    for (size_t k = 0, s = 0; k < nf; ++k) { first[k] = s; s += (size_t)hdr[k].blocksize; }
    {
        flac_job job;
        memset(&job, 0, sizeof(job));
        job.d = d; job.n = n; job.hdr = hdr; job.first = first; job.n_frames = nf; job.nch = f->channels;
        job.out16 = f->bits_per_sample == 16;
        job.pcm = malloc(samples * (size_t)f->channels * (job.out16 ? sizeof(int16_t) : sizeof(int32_t)));
        if (!job.pcm) goto out;
        /* long streams: the device decoder, if the library provides one (BLX_FLAC_GPU=0 turns it off) */
        {
            const char *sw = getenv("BLX_FLAC_GPU"), *mn = getenv("BLX_FLAC_GPU_MIN_SAMPLES");
            const unsigned long long min_samples = mn ? strtoull(mn, NULL, 10) : 1000000ull; /* measured crossover ~0.7 M (tools/flac_gpu_probe.py) */
            if (blx_flac_accel && !(sw && sw[0] == '0') && (unsigned long long)samples * (unsigned)f->channels >= min_samples &&
                sizeof(size_t) == sizeof(uint64_t) &&
                blx_flac_accel(f, d, n, hdr, (const uint64_t *)first, nf, f->channels, job.out16, (uint64_t)samples, job.pcm) == 0) {
                size_t nframes = samples;
                if (total && nframes > total) nframes = (size_t)total;
                if (f->resampled16) free(job.pcm); /* the accelerator resampled on the device as well: nothing decoded comes back */
                else if (job.out16) f->samples16 = (int16_t *)job.pcm;
                else f->samples = (int32_t *)job.pcm;
                f->n_frames = nframes;
                rc = 0;
                goto out;
            }
        }
        pthread_mutex_init(&job.lock, NULL);
        pthread_t th[32];
        int started = 0;
        for (int t = 0; t < threads - 1; ++t)
            if (pthread_create(&th[started], NULL, flac_worker, &job) == 0) started++;
        flac_worker(&job); /* the calling thread works too */
        for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
        pthread_mutex_destroy(&job.lock);
        if (job.failed) { free(job.pcm); goto out; }
        size_t nframes = samples;
        if (total && nframes > total) nframes = (size_t)total;
        if (!nframes) { free(job.pcm); goto out; }
        if (job.out16) f->samples16 = (int16_t *)job.pcm;
        else f->samples = (int32_t *)job.pcm;
        f->n_frames = nframes;
        rc = 0;
    }
out:
    free(hdr);
    free(first);
    return rc;
}

/* The audio frames from d[pos] on. */
static int decode_flac_frames(const uint8_t *d, size_t n, size_t pos, uint64_t total, blx_pcm_file *f) {
    const int threads = flac_decode_threads();
    if (threads > 1 && n - pos >= ((size_t)1 << 18) && decode_flac_parallel(d, n, pos, total, f, threads) == 0) return 0;

    /* sequential: one frame after the other, resynchronising on the next good header after anything that is not a frame */
    /* STREAMINFO's sample count is a 36-bit field of an untrusted file: it sizes the first allocation only as far as the
     * file could plausibly deliver (16 samples per byte); a longer stream grows the buffer as it is decoded. */
    size_t cap = total ? (size_t)total : (size_t)1 << 20;
    if (cap > n * 16 + 65536) cap = n * 16 + 65536;
    const int out16 = f->bits_per_sample == 16;
    const size_t ssize = out16 ? sizeof(int16_t) : sizeof(int32_t);
    void *pcm = malloc(cap * (size_t)f->channels * ssize);
    int32_t *chbuf = (int32_t *)malloc((size_t)65536 * 8 * sizeof(int32_t));
    if (!pcm || !chbuf) { free(pcm); free(chbuf); return -1; }
    size_t nframes = 0;
    const int nch = f->channels;
    while (pos + 2 <= n) {
        flac_hdr h;
        if (parse_frame_header(d, n, pos, f, &h) != 0) { pos++; continue; }
        size_t end = 0;
        const int fr = decode_frame(d, n, &h, chbuf, 65536, &end);
        if (fr < 0) break;
        if (fr > 0) { pos++; continue; } /* damaged frame: resynchronise */
        if (nframes + (size_t)h.blocksize > cap) {
            cap = (nframes + (size_t)h.blocksize) * 2;
            void *np = realloc(pcm, cap * (size_t)nch * ssize);
            if (!np) { free(pcm); free(chbuf); return -1; }
            pcm = np;
        }
        if (out16) interleave16((int16_t *)pcm + nframes * (size_t)nch, chbuf, h.blocksize, nch);
        else interleave((int32_t *)pcm + nframes * (size_t)nch, chbuf, h.blocksize, nch);
        nframes += (size_t)h.blocksize;
        pos = end;
    }
    free(chbuf);
    if (total && nframes > total) nframes = (size_t)total;
    if (!nframes) { free(pcm); return -1; }
    if (out16) f->samples16 = (int16_t *)pcm;
    else f->samples = (int32_t *)pcm;
    f->n_frames = nframes;
    return 0;
}

static int decode_flac(const uint8_t *d, size_t n, blx_pcm_file *f) {
    size_t pos = 4;
    int have_info = 0, last = 0;
    uint64_t total = 0;
    while (!last) {
        if (pos + 4 > n) return -1;
        last = d[pos] >> 7;
        int type = d[pos] & 0x7F;
        size_t len = ((size_t)d[pos + 1] << 16) | ((size_t)d[pos + 2] << 8) | d[pos + 3];
        pos += 4;
        if (pos + len > n) return -1;
        if (type == 0 && len >= 34) {
            const uint8_t *s = d + pos;
            f->sample_rate = (int)(((uint32_t)s[10] << 12) | ((uint32_t)s[11] << 4) | (s[12] >> 4));
            f->channels = ((s[12] >> 1) & 7) + 1;
            f->bits_per_sample = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
            total = ((uint64_t)(s[13] & 0xF) << 32) | ((uint64_t)s[14] << 24) | ((uint64_t)s[15] << 16) |
                    ((uint64_t)s[16] << 8) | s[17];
            memcpy(f->md5, s + 18, 16);
            have_info = 1;
        } else if (type == 4 && len >= 8) { /* VORBIS_COMMENT, little-endian lengths */
            const uint8_t *s = d + pos;
            size_t o = 0;
            uint32_t vlen = s[0] | (s[1] << 8) | (s[2] << 16) | ((uint32_t)s[3] << 24);
            o = 4 + (size_t)vlen;
            if (o + 4 <= len) {
                uint32_t cnt = s[o] | (s[o + 1] << 8) | (s[o + 2] << 16) | ((uint32_t)s[o + 3] << 24);
                o += 4;
                for (uint32_t i = 0; i < cnt && o + 4 <= len; ++i) {
                    uint32_t l = s[o] | (s[o + 1] << 8) | (s[o + 2] << 16) | ((uint32_t)s[o + 3] << 24);
                    o += 4;
                    if (o + l > len) break;
                    add_tag(f, (const char *)s + o, l);
                    o += l;
                }
            }
        }
        pos += len;
    }
    if (!have_info || f->channels < 1 || f->channels > 8) return -1;
    f->is_float = 0;
    f->container = 0;
    return decode_flac_frames(d, n, pos, total, f);
}

/* ------------------------------------------------------------------ */
/* RIFF / WAVE                                                         */
/* ------------------------------------------------------------------ */
static uint32_t le32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

static int decode_wav(const uint8_t *d, size_t n, blx_pcm_file *f) {
    size_t pos = 12;
    int fmt_tag = 0, have_fmt = 0;
    while (pos + 8 <= n) {
        uint32_t len = le32(d + pos + 4);
        const uint8_t *body = d + pos + 8;
        if (pos + 8 + len > n) len = (uint32_t)(n - pos - 8);
        if (!memcmp(d + pos, "fmt ", 4) && len >= 16) {
            fmt_tag = le16(body);
            f->channels = le16(body + 2);
            f->sample_rate = (int)le32(body + 4);
            f->bits_per_sample = le16(body + 14);
            if (fmt_tag == 0xFFFE && len >= 26) fmt_tag = le16(body + 24);
            have_fmt = 1;
        } else if (!memcmp(d + pos, "data", 4) && have_fmt) {
            int bytes = f->bits_per_sample / 8;
            if (bytes < 1 || bytes > 4 || f->channels < 1) return -1;
            size_t total = len / (size_t)bytes;
            size_t nframes = total / (size_t)f->channels;
            f->is_float = (fmt_tag == 3);
            f->container = 1;
            if (f->is_float && bytes != 4) return -1;
            if (bytes == 2 && !f->is_float) { /* the common case: one copy, no widening (a little-endian host is assumed throughout) */
                int16_t *p16 = (int16_t *)malloc((total ? total : 1) * sizeof(int16_t));
                if (!p16) return -1;
                memcpy(p16, body, nframes * (size_t)f->channels * sizeof(int16_t));
                f->samples16 = p16;
                f->n_frames = nframes;
                return nframes ? 0 : -1;
            }
            int32_t *pcm = (int32_t *)malloc((total ? total : 1) * sizeof(int32_t));
            if (!pcm) return -1;
            for (size_t i = 0; i < nframes * (size_t)f->channels; ++i) {
                const uint8_t *s = body + i * (size_t)bytes;
                int32_t v;
                if (bytes == 1) v = (int32_t)s[0] - 128;
                else if (bytes == 2) v = (int16_t)le16(s);
                else if (bytes == 3) v = (int32_t)((uint32_t)s[0] << 8 | (uint32_t)s[1] << 16 | (uint32_t)s[2] << 24) >> 8;
                else v = (int32_t)le32(s); /* s32, or raw f32 bits when is_float */
                pcm[i] = v;
            }
            f->samples = pcm;
            f->n_frames = nframes;
            return nframes ? 0 : -1;
        }
        pos += 8 + (size_t)len + (len & 1);
    }
    return -1;
}

/* ------------------------------------------------------------------ */
int blx_pcm_file_samples32(blx_pcm_file *f) {
    if (f->samples) return 0;
    if (!f->samples16) return -1;
    const size_t n = f->n_frames * (size_t)f->channels;
    int32_t *pcm = (int32_t *)malloc((n ? n : 1) * sizeof(int32_t));
    if (!pcm) return -1;
    for (size_t i = 0; i < n; ++i) pcm[i] = f->samples16[i];
    f->samples = pcm;
    return 0;
}

/* 16-bit PCM WAVE whose "data" chunk starts inside the first `nh` bytes (`head`): the samples are read from the file
 * straight into the buffer the song will own - one copy of the 32 MB of a 3-minute song instead of three.
 * 1 = done, 0 = not this kind of file (the caller reads it whole), -1 = error. */
static int read_wav16_direct(FILE *fp, const uint8_t *head, size_t nh, size_t sz, blx_pcm_file *f) {
    if (nh < 12 || memcmp(head, "RIFF", 4) || memcmp(head + 8, "WAVE", 4)) return 0;
    size_t pos = 12;
    int fmt_tag = 0, have_fmt = 0;
    while (pos + 8 <= nh) {
        const uint32_t len = le32(head + pos + 4);
        if (!memcmp(head + pos, "fmt ", 4) && len >= 16 && pos + 8 + len <= nh) {
            const uint8_t *body = head + pos + 8;
            fmt_tag = le16(body);
            f->channels = le16(body + 2);
            f->sample_rate = (int)le32(body + 4);
            f->bits_per_sample = le16(body + 14);
            if (fmt_tag == 0xFFFE && len >= 26) fmt_tag = le16(body + 24);
            have_fmt = 1;
        } else if (!memcmp(head + pos, "data", 4)) {
            if (!have_fmt || fmt_tag != 1 || f->bits_per_sample != 16 || f->channels < 1) return 0;
            const size_t off = pos + 8;
            size_t bytes = len;
            if (off + bytes > sz) bytes = sz - off;
            const size_t nframes = bytes / 2 / (size_t)f->channels;
            bytes = nframes * (size_t)f->channels * 2;
            if (!nframes) return -1;
            int16_t *p16 = (int16_t *)malloc(bytes);
            if (!p16) return -1;
            const size_t in_head = off < nh ? (nh - off < bytes ? nh - off : bytes) : 0;
            memcpy(p16, head + off, in_head);
            if (fseek(fp, (long)(off + in_head), SEEK_SET) || fread((uint8_t *)p16 + in_head, 1, bytes - in_head, fp) != bytes - in_head) {
                free(p16);
                return -1;
            }
            f->samples16 = p16;
            f->n_frames = nframes;
            f->container = 1;
            return 1;
        }
        pos += 8 + (size_t)len + (len & 1);
    }
    return 0;
}

int blx_pcm_file_read(const char *filename, blx_pcm_file *f) {
    memset(f, 0, sizeof(*f));
    FILE *fp = fopen(filename, "rb");
    if (!fp) return -1;
    fseek(fp, 0, SEEK_END);
    long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    if (sz < 16) { fclose(fp); return -1; }
    f->file_bytes = (uint64_t)sz;
    uint8_t head[4096];
    const size_t nh = (size_t)sz < sizeof(head) ? (size_t)sz : sizeof(head);
    if (fread(head, 1, nh, fp) != nh) { fclose(fp); return -1; }
    const int direct = read_wav16_direct(fp, head, nh, (size_t)sz, f);
    if (direct) {
        fclose(fp);
        if (direct < 0) blx_pcm_file_free(f);
        return direct < 0 ? -1 : 0;
    }
    memset(f, 0, sizeof(*f));
    f->file_bytes = (uint64_t)sz;
    uint8_t *d = (uint8_t *)malloc((size_t)sz);
    if (d) memcpy(d, head, nh);
    if (!d || fread(d + nh, 1, (size_t)sz - nh, fp) != (size_t)sz - nh) { free(d); fclose(fp); return -1; }
    fclose(fp);
    int rc = -1;
    size_t off = 0;
    if (!memcmp(d, "ID3", 3) && sz > 10) /* skip an ID3v2 tag in front of a FLAC stream */
        off = 10 + (((size_t)d[6] & 0x7F) << 21 | ((size_t)d[7] & 0x7F) << 14 | ((size_t)d[8] & 0x7F) << 7 | ((size_t)d[9] & 0x7F));
    if (off + 4 < (size_t)sz && !memcmp(d + off, "fLaC", 4)) rc = decode_flac(d + off, (size_t)sz - off, f);
    else if (!memcmp(d, "RIFF", 4) && !memcmp(d + 8, "WAVE", 4)) rc = decode_wav(d, (size_t)sz, f);
    free(d);
    if (rc) blx_pcm_file_free(f);
    return rc;
}

void blx_pcm_file_free(blx_pcm_file *f) {
    free(f->samples);
    free(f->samples16);
    free(f->resampled16);
    free(f->artist); free(f->title); free(f->album); free(f->tracknumber); free(f->genre);
    memset(f, 0, sizeof(*f));
}
