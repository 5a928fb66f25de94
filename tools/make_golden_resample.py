"""Pins the decode-stage resampler specification (include/blx_resample.h, oracle/resample.c) against the real
libswresample (6.1.100, vendored by opencv-python-headless; build container only) and writes the small fixture
tests/golden/resample_vectors.npz that travels: seeded inputs for several rates / sample formats / channel counts and
the library's int16 / 22 050 Hz / stereo output for each. Also re-checks the reference's md5 pins of its two 48 kHz
fixtures (reference tests/test_decode.c:35-36,55-56) through our FLAC reader + the oracle resampler.
    python tools/make_golden_resample.py
"""
import hashlib
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
from swr_vendored import resample  # noqa: E402

from oracle.binding import Oracle  # noqa: E402

orc = Oracle()
rng = np.random.default_rng(20260101)
cases = {}
n_checked = 0
for rate in (8000, 11025, 16000, 32000, 44100, 48000, 88200, 96000, 22050):
    for kind, fmt, bits in (("s16", "s16", 16), ("s32", "s32", 24), ("f32", "flt", 32), ("u8", "u8", 8)):
        for ch in (1, 2):
            if rate == 22050 and kind == "s16":
                continue  # the reference passes int16 / 22 050 Hz through without libswresample
            n = int(rng.integers(900, 1400))
            if kind == "s16":
                x = (rng.standard_normal((n, ch)) * 9000).clip(-32768, 32767).astype(np.int32)
                x[40:44] = 32767; x[60:63] = -32768
                lib_in, k = x.astype(np.int16), orc.RS_S16
            elif kind == "s32":
                x = (rng.standard_normal((n, ch)) * 0.3 * 2 ** 23).clip(-2 ** 23, 2 ** 23 - 1).astype(np.int32)
                x[40:44] = 2 ** 23 - 1; x[60:63] = -2 ** 23
                lib_in, k = (x << 8).astype(np.int32), orc.RS_S32  # FFmpeg's decoders left-justify 24-bit samples
            elif kind == "f32":
                f = (rng.standard_normal((n, ch)) * 0.4).astype(np.float32)
                f[40:44] = 1.5; f[60:63] = -2.0  # beyond full scale: clipped on output
                x, lib_in, k = f.view(np.int32), f, orc.RS_F32
            else:
                u = rng.integers(0, 256, (n, ch)).astype(np.uint8)
                x, lib_in, k = u.astype(np.int32) - 128, u, orc.RS_U8
            want = resample(lib_in.reshape(-1), ch, rate, fmt, "s16")
            got = orc.resample_to_s16(x.reshape(-1), k, bits, ch, rate)
            assert got.shape == want.shape and np.array_equal(got, want), (rate, kind, ch, got.shape, want.shape)
            n_checked += 1
            if rate in (8000, 44100, 48000, 96000, 22050) and (ch == 2 or kind == "s32" or rate == 8000):
                key = f"{kind}_{bits}_{ch}_{rate}"
                cases[key + "_in"] = x.reshape(-1).astype(np.int32)
                cases[key + "_out"] = want
print("oracle == libswresample on", n_checked, "cases;", len(cases) // 2, "kept as fixtures")
# long signals: block boundaries of the library's streaming interface
for rate, n in ((48000, 200003), (44100, 131072)):
    x = (rng.standard_normal((n, 2)) * 9000).clip(-32768, 32767).astype(np.int32)
    assert np.array_equal(orc.resample_to_s16(x.reshape(-1), orc.RS_S16, 16, 2, rate),
                          resample(x.astype(np.int16).reshape(-1), 2, rate, "s16", "s16"))
np.savez_compressed(os.path.join(REPO, "tests", "golden", "resample_vectors.npz"), **cases)

# the reference's own pins, through our FLAC reader (host C) + the oracle
if os.path.isdir("/root/reference/audio"):
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from flac_util import read_pcm_file
    for fn, pin in (("song_s32.flac", "eb9f31a7b9ed022d66ff82b76e7c3c18"), ("song_s32_mono.flac", "747dbfcd75bebc23ebe2024935aede36")):
        a, n, ch, rate, bps = read_pcm_file("/root/reference/audio/" + fn)
        pcm = orc.resample_to_s16(a, orc.RS_S32, bps, ch, rate)
        md5 = hashlib.md5(pcm.tobytes()).hexdigest()
        print(fn, n, ch, rate, bps, "->", len(pcm), md5, "PIN OK" if md5 == pin else "PIN MISMATCH")
        assert md5 == pin
