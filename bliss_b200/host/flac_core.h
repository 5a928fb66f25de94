/* flac_core.h — the part of the FLAC decoder that works on one frame: bit reader, residual, subframes, stereo
 * decorrelation (RFC 9639). Shared by the host reader (flac_reader.c) and the device decoder (csrc/flacdec.cu): plain
 * functions on byte buffers, no allocation, no library calls. The includer provides
 *   FLAC_LPC(out, n, order, coef, shift, narrow)   LPC synthesis (default: fc_lpc_generic below)
 *   FLAC_CRC16(p, n)                                CRC-16 of a byte range
 * Written from the FLAC format specification; not derived from libFLAC or FFmpeg source. */
#ifndef BLX_FLAC_CORE_H
#define BLX_FLAC_CORE_H
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifndef FLAC_FN /* (a device-only includer defines it as `__device__ static inline`) */
#ifdef __CUDACC__
#define FLAC_FN __host__ __device__ static inline
#else
#define FLAC_FN static inline
#endif
#endif
#ifdef __CUDA_ARCH__
#define FLAC_CLZ64(x) __clzll((long long)(x))
#else
#define FLAC_CLZ64(x) __builtin_clzll(x)
#endif

/* ------------------------------------------------------------------ */
/* bit reader                                                          */
/* ------------------------------------------------------------------ */
typedef struct {
    const uint8_t *p;
    size_t n;
    size_t pos;   /* next byte to load (runs past n at the end of the data: zeros are shifted in) */
    uint64_t acc; /* bit window, MSB first: the next bit of the stream is bit 63; bits below the valid `cnt` are 0 */
    int cnt;      /* valid bits in acc */
} bitrd;

FLAC_FN void br_init(bitrd *b, const uint8_t *p, size_t n, size_t pos) {
    b->p = p; b->n = n; b->pos = pos; b->acc = 0; b->cnt = 0;
}

/* Tops the window up to at least 57 valid bits: one unaligned 8-byte load when the data allows, byte by byte at its end. */
FLAC_FN void br_refill(bitrd *b) {
#ifndef __CUDACC__ /* host compilers only: one unaligned 8-byte load */
    if (b->pos + 8 <= b->n) {
        uint64_t w;
        memcpy(&w, b->p + b->pos, 8);
        w = __builtin_bswap64(w); /* little-endian host (as everywhere in this library) */
        b->acc |= w >> b->cnt;
        const int take = (63 - b->cnt) >> 3; /* whole bytes that fit */
        b->pos += (size_t)take;
        b->cnt += take * 8;
        b->acc &= ~(~0ull >> b->cnt); /* the bits of the partly loaded next byte stay out of the window (56 <= cnt <= 63) */
        return;
    }
#endif
#ifdef __CUDA_ARCH__
    /* device: the lanes of a warp read different streams, so a data-dependent refill loop would make every lane wait for
     * the slowest one at every sample; instead ONE step of four bytes (independent loads, one latency) whenever 32 bits
     * or fewer are left - enough for any single read - and the byte loop only at the very end of the data */
    if (b->cnt > 32) return;
    if (b->pos + 4 <= b->n) {
        const uint8_t *q = b->p + b->pos;
        const uint32_t w = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
        b->acc |= (uint64_t)w << (32 - b->cnt);
        b->cnt += 32;
        b->pos += 4;
        return;
    }
#endif
    while (b->cnt <= 56) {
        const uint64_t byte = (b->pos < b->n) ? b->p[b->pos] : 0;
        b->pos++;
        b->acc |= byte << (56 - b->cnt);
        b->cnt += 8;
    }
}

/* bytes of the stream consumed so far; more than n = the reader ran off the end of the data */
FLAC_FN size_t br_bytepos(const bitrd *b) { return b->pos - (size_t)(b->cnt >> 3); }
FLAC_FN int br_err(const bitrd *b) { return br_bytepos(b) > b->n; }

FLAC_FN uint32_t br_u(bitrd *b, int nbits) { /* nbits <= 32 */
    if (nbits == 0) return 0;
    if (b->cnt < nbits) br_refill(b);
    const uint32_t v = (uint32_t)(b->acc >> (64 - nbits));
    b->acc <<= nbits;
    b->cnt -= nbits;
    return v;
}

FLAC_FN int32_t br_s(bitrd *b, int nbits) {
    uint32_t v = br_u(b, nbits);
    if (nbits == 0) return 0;
    if (nbits < 32 && (v >> (nbits - 1))) v |= ~((1u << nbits) - 1u);
    return (int32_t)v;
}

FLAC_FN uint32_t br_unary(bitrd *b) { /* count zeros before the next 1 bit */
    uint32_t z = 0;
    for (;;) {
        if (b->acc == 0) { /* only zeros in the window */
            z += (uint32_t)b->cnt;
            b->cnt = 0;
            if (b->pos >= b->n + 8) return z; /* nothing but the padding behind the data: the caller sees br_err */
            br_refill(b);
            continue;
        }
        const int lz = FLAC_CLZ64(b->acc); /* < cnt: the bits below cnt are zero */
        z += (uint32_t)lz;
        b->acc <<= lz;   /* two shifts: lz + 1 may be 64 */
        b->acc <<= 1;
        b->cnt -= lz + 1;
        return z;
    }
}

FLAC_FN void br_align(bitrd *b) {
    const int drop = b->cnt & 7;
    b->acc <<= drop;
    b->cnt -= drop;
}


FLAC_FN int fc_ilog2_ceil(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

/* LPC synthesis out[i] += (sum_j coef[j] out[i - 1 - j]) >> shift, i = order..n-1 (RFC 9639 section 9.2.6), plain form.
 * `narrow`: bps + precision + log2(order) <= 32, so 32-bit (unsigned: wrap-around, never signed overflow) words suffice. */
FLAC_FN void fc_lpc_generic(int32_t *out, int n, int order, const int32_t *coef, int shift, int narrow) {
    if (narrow) {
        for (int i = order; i < n; ++i) {
            uint32_t acc = 0;
            for (int j = 0; j < order; ++j) acc += (uint32_t)coef[j] * (uint32_t)out[i - 1 - j];
            out[i] = (int32_t)((uint32_t)((int32_t)acc >> shift) + (uint32_t)out[i]);
        }
    } else {
        for (int i = order; i < n; ++i) {
            int64_t acc = 0;
            for (int j = 0; j < order; ++j) acc += (int64_t)coef[j] * out[i - 1 - j];
            out[i] = (int32_t)((acc >> shift) + out[i]);
        }
    }
}
/* (the 32-bit form works on unsigned words, so a damaged file whose samples outgrow their declared width wraps around
 * instead of running into signed overflow) */
#define FC_LPC_BODY(ACC_T, ORD)                                                               \
    for (int i = order; i < n; ++i) {                                                          \
        ACC_T acc = 0;                                                                         \
        const int32_t *h = out + i - (ORD);                                                    \
        for (int k = 0; k < (ORD); ++k) acc += (ACC_T)rc[k] * (ACC_T)h[k];                     \
        if (sizeof(ACC_T) == 4) out[i] = (int32_t)((uint32_t)((int32_t)acc >> shift) + (uint32_t)out[i]); \
        else out[i] = (int32_t)(((int64_t)acc >> shift) + out[i]);                             \
    }
#define FC_LPC_SWITCH(ACC_T)                                                                  \
    switch (order) {                                                                           \
        case 1: FC_LPC_BODY(ACC_T, 1) break;   case 2: FC_LPC_BODY(ACC_T, 2) break;          \
        case 3: FC_LPC_BODY(ACC_T, 3) break;   case 4: FC_LPC_BODY(ACC_T, 4) break;          \
        case 5: FC_LPC_BODY(ACC_T, 5) break;   case 6: FC_LPC_BODY(ACC_T, 6) break;          \
        case 7: FC_LPC_BODY(ACC_T, 7) break;   case 8: FC_LPC_BODY(ACC_T, 8) break;          \
        case 9: FC_LPC_BODY(ACC_T, 9) break;   case 10: FC_LPC_BODY(ACC_T, 10) break;        \
        case 11: FC_LPC_BODY(ACC_T, 11) break; case 12: FC_LPC_BODY(ACC_T, 12) break;        \
        default: FC_LPC_BODY(ACC_T, order) break;                                             \
    }


/* The taps reversed once (the dot product then runs over contiguous samples), a loop with a constant trip count for the
 * common orders (unrolled by the compiler: independent loads, a multiply-add tree), the plain loop for the rest. */
FLAC_FN void fc_lpc_unrolled(int32_t *out, int n, int order, const int32_t *coef, int shift, int narrow) {
    int32_t rc[32];
    for (int k = 0; k < order; ++k) rc[k] = coef[order - 1 - k];
    if (narrow) { FC_LPC_SWITCH(uint32_t) }
    else { FC_LPC_SWITCH(int64_t) }
}
#ifndef FLAC_LPC
#define FLAC_LPC fc_lpc_unrolled
#endif

#ifndef BLX_RICE_FAST
#define BLX_RICE_FAST 1
#endif
FLAC_FN int read_residual(bitrd *b, int32_t *out, int blocksize, int pred_order) {
    int method = (int)br_u(b, 2);
    if (method > 1) return -1;
    int pbits = method ? 5 : 4;
    uint32_t escape = method ? 31u : 15u;
    int porder = (int)br_u(b, 4);
    int nparts = 1 << porder;
    if ((blocksize >> porder) << porder != blocksize && porder > 0) return -1;
    int idx = pred_order;
    for (int part = 0; part < nparts; ++part) {
        int count = (blocksize >> porder) - (part == 0 ? pred_order : 0);
        if (count < 0) return -1;
        uint32_t param = br_u(b, pbits);
        if (param == escape) {
            int raw = (int)br_u(b, 5);
            for (int i = 0; i < count; ++i) out[idx++] = br_s(b, raw);
        } else {
            for (int i = 0; i < count; ++i) {
                uint32_t u;
                if (b->cnt < 57) br_refill(b);
                const int lz = b->acc ? FLAC_CLZ64(b->acc) : 64;
                if (BLX_RICE_FAST && param <= 16 && lz + 1 + (int)param <= b->cnt) {
                    /* the usual case in one go: the zeros, the stop bit and the low bits all sit in the window */
                    const uint64_t rest = b->acc << lz << 1;
                    const uint32_t r = (uint32_t)((rest >> 32) >> (32 - param)); /* param = 0: a 32-bit word shifted out whole */
                    u = ((uint32_t)lz << param) | (param ? r : 0u);
                    b->acc = rest << param;
                    b->cnt -= lz + 1 + (int)param;
                } else {
                    const uint32_t q = br_unary(b);
                    const uint32_t r = br_u(b, (int)param);
                    u = (q << param) | r;
                }
                out[idx++] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
            }
        }
        if (br_err(b)) return -1;
    }
    return 0;
}

FLAC_FN int read_subframe(bitrd *b, int32_t *out, int blocksize, int bps) {
    if (br_u(b, 1)) return -1; /* padding */
    int type = (int)br_u(b, 6);
    int wasted = 0;
    if (br_u(b, 1)) wasted = (int)br_unary(b) + 1;
    bps -= wasted;
    /* samples are kept in int32: the 33-bit side channel of a 32-bit stereo stream is not supported */
    if (bps <= 0 || bps > 32) return -1;
    if (type == 0) { /* constant */
        int32_t v = br_s(b, bps);
        for (int i = 0; i < blocksize; ++i) out[i] = v;
    } else if (type == 1) { /* verbatim */
        for (int i = 0; i < blocksize; ++i) out[i] = br_s(b, bps);
    } else if (type >= 8 && type <= 12) { /* fixed predictor */
        int order = type - 8;
        if (order > blocksize) return -1;
        for (int i = 0; i < order; ++i) out[i] = br_s(b, bps);
        if (read_residual(b, out, blocksize, order)) return -1;
        for (int i = order; i < blocksize; ++i) {
            int64_t p = 0;
            switch (order) {
                case 1: p = out[i - 1]; break;
                case 2: p = 2 * (int64_t)out[i - 1] - out[i - 2]; break;
                case 3: p = 3 * (int64_t)out[i - 1] - 3 * (int64_t)out[i - 2] + out[i - 3]; break;
                case 4: p = 4 * (int64_t)out[i - 1] - 6 * (int64_t)out[i - 2] + 4 * (int64_t)out[i - 3] - out[i - 4]; break;
                default: break;
            }
            out[i] = (int32_t)(p + out[i]);
        }
    } else if (type >= 32) { /* LPC */
        int order = type - 31;
        if (order > blocksize) return -1;
        int32_t coef[32];
        for (int i = 0; i < order; ++i) out[i] = br_s(b, bps);
        int prec = (int)br_u(b, 4) + 1;
        if (prec == 16) return -1;
        int shift = br_s(b, 5);
        if (shift < 0) return -1;
        for (int i = 0; i < order; ++i) coef[i] = br_s(b, prec);
        if (read_residual(b, out, blocksize, order)) return -1;
        FLAC_LPC(out, blocksize, order, coef, shift, bps + prec + fc_ilog2_ceil(order) <= 32);
    } else {
        return -1;
    }
    if (wasted)
        for (int i = 0; i < blocksize; ++i) out[i] = (int32_t)((uint32_t)out[i] << wasted);
    return br_err(b) ? -1 : 0;
}

/* ---- frames ---------------------------------------------------------------------------------------------------- */
typedef struct {
    size_t off;        /* byte offset of the sync code */
    size_t hdr_len;    /* header bytes incl. CRC-8 */
    int blocksize, ch_code, bps;
    int variable;      /* blocking strategy bit: the coded number counts samples, not frames */
    uint64_t number;   /* coded frame / sample number */
} flac_hdr;

/* Decodes the frame whose (checked) header is h into chbuf (channel c at chbuf + c * ch_stride), stereo decorrelation
 * undone; *end = the byte behind its CRC-16. 0 = good frame, 1 = damaged (resynchronise), -1 = ran off the data. */
FLAC_FN int decode_frame(const uint8_t *d, size_t n, const flac_hdr *h, int32_t *chbuf, size_t ch_stride, size_t *end) {
    bitrd b;
    br_init(&b, d, n, h->off + h->hdr_len);
    const int nch = (h->ch_code < 8) ? h->ch_code + 1 : 2, blocksize = h->blocksize;
    for (int c = 0; c < nch; ++c) {
        const int side = (h->ch_code == 8 && c == 1) || (h->ch_code == 9 && c == 0) || (h->ch_code == 10 && c == 1);
        if (read_subframe(&b, chbuf + (size_t)c * ch_stride, blocksize, h->bps + side)) return 1;
    }
    br_align(&b);
    const size_t body_end = br_bytepos(&b);
    const uint32_t crc16 = br_u(&b, 16);
    if (br_err(&b)) return -1;
    if (crc16 != FLAC_CRC16(d + h->off, body_end - h->off)) return 1;
    int32_t *c0 = chbuf, *c1 = chbuf + ch_stride;
    if (h->ch_code == 8) { for (int i = 0; i < blocksize; ++i) c1[i] = c0[i] - c1[i]; }
    else if (h->ch_code == 9) { for (int i = 0; i < blocksize; ++i) c0[i] = c0[i] + c1[i]; }
    else if (h->ch_code == 10) {
        for (int i = 0; i < blocksize; ++i) {
            const int32_t side = c1[i];
            const int32_t mid = (int32_t)(((uint32_t)c0[i] << 1) | (uint32_t)(side & 1));
            c0[i] = (mid + side) >> 1;
            c1[i] = (mid - side) >> 1;
        }
    }
    *end = br_bytepos(&b);
    return 0;
}

#endif /* BLX_FLAC_CORE_H */
