"""How far is the engine's 44.1 kHz float32 front-end (include/blx_frontend.h: 23-tap half-band 2:1 decimation +
int16 rounding) from what the reference's decoder would hand the analysers for the same file (libswresample,
reference src/decode.c:323-345: float/44.1 kHz -> s16/22 050 Hz/stereo)?

Round 2: FILES no longer take this front-end - bl_audio_decode runs the libswresample-exact resampler of
include/blx_resample.h (deviation 0, md5 pins reproduced), and blx_analyze_batch_f32_exact does the same for float
buffers a caller holds. What this script measures is therefore only the default batch ABI for raw 44.1 kHz float
streams (blx_analyze_batch_f32, the benchmark's workload), whose cheaper half-band filter is a documented choice.

Runs in the build container only: drives the libswresample vendored in opencv's wheel through ctypes (see
tools/make_golden_s32.py) and the oracle's front-end + analysers on CPU. Prints sample-level differences
and the resulting force-vector differences for a few synthetic songs.
"""
import ctypes, glob, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from oracle.binding import Oracle
from synth import song_f32
LIBS = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs"
for name in ("libdrm", "libcrypto", "libssl"):
    for p in glob.glob(os.path.join(LIBS, name + "*")):
        try: ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
        except OSError: pass
avutil = ctypes.CDLL(glob.glob(os.path.join(LIBS, "libavutil-*"))[0], mode=ctypes.RTLD_GLOBAL)
swr = ctypes.CDLL(glob.glob(os.path.join(LIBS, "libswresample-*"))[0], mode=ctypes.RTLD_GLOBAL)
swr.swr_alloc.restype = ctypes.c_void_p
avutil.av_opt_set.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
avutil.av_opt_set_int.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int]
swr.swr_init.argtypes = [ctypes.c_void_p]
swr.swr_convert.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
AV_FLT, AV_S16 = 3, 1


def swresample_f32_mono(x):
    ctx = swr.swr_alloc()
    assert avutil.av_opt_set(ctx, b"in_chlayout", b"mono", 0) == 0 and avutil.av_opt_set(ctx, b"out_chlayout", b"stereo", 0) == 0
    for k, v in ((b"in_sample_rate", 44100), (b"out_sample_rate", 22050), (b"in_sample_fmt", AV_FLT), (b"out_sample_fmt", AV_S16)):
        assert avutil.av_opt_set_int(ctx, k, v, 0) == 0
    assert swr.swr_init(ctx) == 0
    out, BLK = [], 4096
    obuf = np.zeros(2 * (BLK + 4096), dtype=np.int16)
    def conv(inp, n_in):
        op = (ctypes.c_void_p * 1)(obuf.ctypes.data)
        ip = (ctypes.c_void_p * 1)(inp.ctypes.data) if inp is not None else None
        got = swr.swr_convert(ctx, op, BLK + 4096, ip, n_in)
        if got > 0: out.append(obuf[:2 * got].copy())
        return got
    for s in range(0, len(x), BLK):
        blk = np.ascontiguousarray(x[s:s + BLK])
        conv(blk, len(blk))
    while conv(None, 0) > 0:
        pass
    return np.concatenate(out)


def music_like(x):
    """-12 dB/octave above ~2 kHz (two one-pole low-passes): the spectrum of the test songs is nearly white up to
    22 kHz, real music is not."""
    from scipy.signal import lfilter
    a = float(np.exp(-2 * np.pi * 2000.0 / 44100.0))
    y = lfilter([1 - a], [1, -a], lfilter([1 - a], [1, -a], x.astype(np.float64)))
    return (y * (0.25 / max(np.max(np.abs(y)), 1e-9))).astype(np.float32)


orc = Oracle()
for seed, sec, shape in ((1, 8.0, "white"), (2, 15.0, "white"), (3, 30.0, "white"), (1, 8.0, "music-like"), (3, 30.0, "music-like")):
    x = song_f32(900 + seed, sec)
    if shape != "white":
        x = music_like(x)
    print(f"[{shape} spectrum]")
    ours = orc.frontend_f32(x)                      # interleaved L = R
    ref = swresample_f32_mono(x)
    n = min(len(ours), len(ref))
    # swresample's filter has its own group delay compensation; align by the best lag within +-4 samples
    a, best = ours[0:n:2].astype(np.int32), None
    for lag in range(-4, 5):
        b = ref[0:n:2].astype(np.int32)
        d = a[max(0, lag):len(a) + min(0, lag)] - b[max(0, -lag):len(b) + min(0, -lag)]
        rms = float(np.sqrt(np.mean(d[64:-64].astype(np.float64) ** 2)))
        if best is None or rms < best[1]: best = (lag, rms, int(np.max(np.abs(d[64:-64]))))
    r1, r2 = orc.analyze(ours, int(sec)), orc.analyze(ref, int(sec))
    print(f"song {seed} ({sec:g} s): lag {best[0]}, rms diff {best[1]:.2f} LSB, max |diff| {best[2]} LSB; signal rms {np.sqrt(np.mean(a.astype(np.float64) ** 2)):.0f} LSB")
    for k in ("tempo", "amplitude", "frequency", "attack"):
        print(f"    {k:10s} ours {r1[k]:+.6f}  swresample {r2[k]:+.6f}  diff {r1[k] - r2[k]:+.2e}")
    print(f"    beat ours {r1['beat']} swresample {r2['beat']}")
