"""Host-side file reader of the product (bliss_b200/host/flac_reader.c: FLAC + RIFF/WAVE, plain C, no GPU) for tests
and fixture generators."""
import ctypes

import numpy as np

import bliss_b200


class PcmFile(ctypes.Structure):
    _fields_ = [("samples", ctypes.POINTER(ctypes.c_int32)), ("samples16", ctypes.POINTER(ctypes.c_int16)), ("n_frames", ctypes.c_size_t), ("channels", ctypes.c_int),
                ("sample_rate", ctypes.c_int), ("bits_per_sample", ctypes.c_int), ("is_float", ctypes.c_int),
                ("container", ctypes.c_int), ("file_bytes", ctypes.c_uint64), ("md5", ctypes.c_uint8 * 16)] + [
                    (k, ctypes.c_char_p) for k in ("artist", "title", "album", "tracknumber", "genre")] + [
                    ("resampled16", ctypes.POINTER(ctypes.c_int16)), ("resampled_frames", ctypes.c_size_t)]


def read_pcm_file(path):
    """(int32 samples interleaved, n_frames, channels, sample_rate, bits_per_sample)"""
    L = ctypes.CDLL(bliss_b200.LIB_PATH)
    L.blx_pcm_file_read.argtypes = [ctypes.c_char_p, ctypes.POINTER(PcmFile)]
    L.blx_pcm_file_samples32.argtypes = [ctypes.POINTER(PcmFile)]
    L.blx_pcm_file_free.argtypes = [ctypes.POINTER(PcmFile)]
    f = PcmFile()
    if L.blx_pcm_file_read(str(path).encode(), ctypes.byref(f)) != 0:
        raise IOError(f"cannot read {path}")
    if L.blx_pcm_file_samples32(ctypes.byref(f)) != 0:
        raise MemoryError
    a = np.ctypeslib.as_array(f.samples, (f.n_frames * f.channels,)).copy()
    out = (a, int(f.n_frames), int(f.channels), int(f.sample_rate), int(f.bits_per_sample))
    L.blx_pcm_file_free(ctypes.byref(f))
    return out


def repeat_flac(path, times):
    """A long FLAC stream made of `times` copies of the audio frames of a (fixed-blocksize, libFLAC-style) file, renumbered and
    re-checksummed: decodes to the original PCM repeated. For tests and probes that need minutes of audio without an
    encoder run."""
    from flac_encode import _crc, _utf8
    d = open(path, "rb").read()
    assert d[:4] == b"fLaC"
    pos, last = 4, 0
    while not last:
        last = d[pos] >> 7
        ln = int.from_bytes(d[pos + 1:pos + 4], "big")
        if d[pos] & 0x7F == 0:
            info_at = pos + 4
        pos += 4 + ln
    audio0 = pos
    # frame starts: sync code, sane header, CRC-8, consecutive frame numbers
    starts, want = [], 0
    i = audio0
    while i + 6 < len(d):
        i = d.find(b"\xff\xf8", i)
        if i < 0:
            break
        num, extra = d[i + 4], 0
        if num & 0x80:
            extra = 1 if num < 0xE0 else 2 if num < 0xF0 else 3
            v = num & (0x3F >> extra) if extra else num
            for k in range(extra):
                v = (v << 6) | (d[i + 5 + k] & 0x3F)
            num = v
        hl = 4 + 1 + extra
        bs_code, sr_code = d[i + 2] >> 4, d[i + 2] & 0xF
        hl += 1 if bs_code == 6 else 2 if bs_code == 7 else 0
        hl += 1 if sr_code == 12 else 2 if sr_code in (13, 14) else 0
        if num == want and _crc(d[i:i + hl], 0x07, 8) == d[i + hl]:
            starts.append((i, hl, extra))
            want += 1
            i += hl
        else:
            i += 1
    assert len(starts) >= 2
    tab = [0] * 256
    for b in range(256):
        c = b << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x8005) & 0xFFFF if c & 0x8000 else (c << 1) & 0xFFFF
        tab[b] = c
    out = bytearray(d[:audio0])
    n_frames_audio = int.from_bytes(d[info_at + 13:info_at + 18], "big") & ((1 << 36) - 1)
    total = n_frames_audio * times
    out[info_at + 13] = (out[info_at + 13] & 0xF0) | ((total >> 32) & 0xF)
    out[info_at + 14:info_at + 18] = (total & 0xFFFFFFFF).to_bytes(4, "big")
    out[info_at + 18:info_at + 34] = bytes(16)  # the md5 no longer applies
    k = 0
    for _ in range(times):
        for j, (off, hl, extra) in enumerate(starts):
            end = starts[j + 1][0] if j + 1 < len(starts) else len(d)
            hdr = d[off:off + 4] + _utf8(k) + d[off + 5 + extra:off + hl]
            hdr += bytes([_crc(hdr, 0x07, 8)])
            frame = hdr + d[off + hl + 1:end - 2]
            c = 0
            for b in frame:
                c = ((c << 8) & 0xFFFF) ^ tab[(c >> 8) ^ b]
            out += frame + c.to_bytes(2, "big")
            k += 1
    return bytes(out)
