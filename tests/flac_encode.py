"""A small FLAC ENCODER for tests (RFC 9639): builds streams that exercise the corners of the host decoder
(bliss_b200/host/flac_reader.c) that the reference's three fixtures do not reach - every subframe type, LPC orders up to
32, wasted bits, escaped Rice partitions, both Rice parameter widths, all stereo decorrelation modes, odd block sizes,
8 / 12 / 16 / 20 / 24-bit samples. Not an efficient encoder: predictors are arbitrary, only bit-exact reversibility matters."""
import struct

import numpy as np


def _crc(data, poly, bits):
    c = 0
    top = 1 << (bits - 1)
    mask = (1 << bits) - 1
    for b in data:
        c ^= b << (bits - 8)
        for _ in range(8):
            c = ((c << 1) ^ poly) & mask if c & top else (c << 1) & mask
    return c


class _Bits:
    def __init__(self):
        self.acc = 0
        self.n = 0

    def u(self, v, nbits):
        if nbits:
            self.acc = (self.acc << nbits) | (int(v) & ((1 << nbits) - 1))
            self.n += nbits

    def s(self, v, nbits):
        self.u(int(v) & ((1 << nbits) - 1), nbits)

    def unary(self, q):
        self.u(0, q)
        self.u(1, 1)

    def align(self):
        if self.n % 8:
            self.u(0, 8 - self.n % 8)

    def bytes(self):
        assert self.n % 8 == 0
        return self.acc.to_bytes(self.n // 8, "big") if self.n else b""


def _utf8(v):
    if v < 0x80:
        return bytes([v])
    out = []
    n = 1
    while v >= (1 << (6 * n + (6 - n))):
        n += 1
    for i in range(n):
        out.append(0x80 | ((v >> (6 * i)) & 0x3F))
    lead = ((0xFF << (7 - n)) & 0xFF) | (v >> (6 * n))
    return bytes([lead] + out[::-1])


def _residual(bw, res, pred_order, blocksize, method, porder, escape_parts=()):
    bw.u(method, 2)
    bw.u(porder, 4)
    pbits = 5 if method else 4
    nparts = 1 << porder
    idx = 0
    for part in range(nparts):
        count = (blocksize >> porder) - (pred_order if part == 0 else 0)
        chunk = [int(x) for x in res[idx:idx + count]]
        idx += count
        if part in escape_parts or not chunk:
            bw.u((1 << pbits) - 1, pbits)
            width = max([1] + [max(x, -x - 1).bit_length() + 1 for x in chunk])
            bw.u(width, 5)
            for x in chunk:
                bw.s(x, width)
            continue
        folded = [(x << 1) ^ (x >> 63) if x >= 0 else ((-x) << 1) - 1 for x in chunk]
        mean = max(1, sum(folded) // len(folded))
        param = min(max(mean.bit_length() - 1, 0), (1 << pbits) - 2)
        bw.u(param, pbits)
        for u in folded:
            bw.unary(u >> param)
            bw.u(u & ((1 << param) - 1), param)
    assert idx == len(res)


def _subframe(bw, x, bps, kind, rng, lpc_order=8, method=0, porder=0, escape_parts=(), wasted=0):
    """x: int64 array of one channel's block, already shifted right by `wasted`."""
    n = len(x)
    bw.u(0, 1)
    types = {"constant": 0, "verbatim": 1}
    if kind in types:
        bw.u(types[kind], 6)
    elif kind.startswith("fixed"):
        order = int(kind[5:])
        bw.u(8 + order, 6)
    else:
        order = lpc_order
        bw.u(31 + order, 6)
    if wasted:
        bw.u(1, 1)
        bw.unary(wasted - 1)
    else:
        bw.u(0, 1)
    w = bps - wasted
    if kind == "constant":
        bw.s(x[0], w)
    elif kind == "verbatim":
        for v in x:
            bw.s(v, w)
    elif kind.startswith("fixed"):
        for v in x[:order]:
            bw.s(v, w)
        coefs = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}[order]
        res = [int(x[i]) - sum(c * int(x[i - 1 - j]) for j, c in enumerate(coefs)) for i in range(order, n)]
        _residual(bw, res, order, n, method, porder, escape_parts)
    else:
        for v in x[:order]:
            bw.s(v, w)
        prec = int(rng.integers(5, 16))
        shift = int(rng.integers(max(0, prec - 6), prec))
        lim = 1 << (prec - 1)
        bound = max(1, (1 << shift) // max(order // 2, 1))  # keeps |prediction| within ~3x the signal: residuals stay small
        coefs = [int(np.clip(c, -lim, lim - 1)) for c in rng.integers(-bound, bound + 1, order)]
        coefs[0] = min(lim - 1, (1 << shift))  # roughly "previous sample"
        bw.u(prec - 1, 4)
        bw.s(shift, 5)
        for c in coefs:
            bw.s(c, prec)
        res = [int(x[i]) - (sum(c * int(x[i - 1 - j]) for j, c in enumerate(coefs)) >> shift) for i in range(order, n)]
        _residual(bw, res, order, n, method, porder, escape_parts)


def encode(pcm, bps, rate, blocksize, plan, seed=0):
    """pcm: int array [frames, channels]; plan(frame_index) -> dict(kind=..., stereo=0|8|9|10, lpc_order=, method=, porder=,
    escape_parts=, wasted=) per frame. Returns the bytes of a FLAC file."""
    rng = np.random.default_rng(seed)
    pcm = np.asarray(pcm, dtype=np.int64)
    n, ch = pcm.shape
    out = bytearray(b"fLaC")
    info = _Bits()
    info.u(blocksize, 16); info.u(blocksize, 16); info.u(0, 24); info.u(0, 24)
    info.u(rate, 20); info.u(ch - 1, 3); info.u(bps - 1, 5); info.u(n, 36)
    out += bytes([0x80]) + (34).to_bytes(3, "big") + info.bytes() + bytes(16)
    ss_code = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6, 32: 7}[bps]
    for fi, start in enumerate(range(0, n, blocksize)):
        blk = pcm[start:start + blocksize]
        bs = len(blk)
        p = dict(kind="lpc", stereo=None, lpc_order=8, method=0, porder=0, escape_parts=(), wasted=0)
        p.update(plan(fi))
        if p["porder"] and (bs >> p["porder"]) << p["porder"] != bs:
            p["porder"] = 0
        order = int(p["kind"][5:]) if p["kind"].startswith("fixed") else (p["lpc_order"] if p["kind"] == "lpc" else 0)
        if order > bs or (bs >> p["porder"]) < order:
            p["kind"], p["porder"] = "verbatim", 0
        hdr = _Bits()
        hdr.u(0xFFF8 >> 2, 14); hdr.u(0, 1); hdr.u(0, 1)
        hdr.u(7, 4); hdr.u(0, 4)
        ch_code = p["stereo"] if (ch == 2 and p["stereo"]) else ch - 1
        hdr.u(ch_code, 4); hdr.u(ss_code, 3); hdr.u(0, 1)
        hb = hdr.bytes() + _utf8(fi) + (bs - 1).to_bytes(2, "big")
        hb += bytes([_crc(hb, 0x07, 8)])
        body = _Bits()
        chans = [blk[:, c] for c in range(ch)]
        widths = [bps] * ch
        if ch == 2 and ch_code == 8:    # left / side
            chans = [chans[0], chans[0] - chans[1]]; widths = [bps, bps + 1]
        elif ch == 2 and ch_code == 9:  # side / right
            chans = [chans[0] - chans[1], chans[1]]; widths = [bps + 1, bps]
        elif ch == 2 and ch_code == 10:  # mid / side
            chans = [(chans[0] + chans[1]) >> 1, chans[0] - chans[1]]; widths = [bps, bps + 1]
        for x, w in zip(chans, widths):
            kind = p["kind"]
            wasted = p["wasted"]
            if kind == "constant" and np.any(x != x[0]):
                kind = "verbatim"
            if wasted and np.any(x & ((1 << wasted) - 1)):
                wasted = 0
            _subframe(body, x >> wasted, w, kind, rng, p["lpc_order"], p["method"], p["porder"], p["escape_parts"], wasted)
        body.align()
        frame = hb + body.bytes()
        out += frame + _crc(frame, 0x8005, 16).to_bytes(2, "big")
    return bytes(out)
