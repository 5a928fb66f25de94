"""Generates tests/golden/song_s32_pcm.npz: the decoded + resampled PCM of the reference's second fixture.

reference audio/song_s32.flac is 48 kHz / 24-bit; bl_audio_decode takes it through libswresample to
int16 / 22 050 Hz / stereo (reference src/decode.c:323-345,388-392). FFmpeg is not in this image, but
opencv's wheel vendors libswresample + libavutil; this script decodes the FLAC with the repo's own reader
(bliss_b200/host/flac_reader.c), left-justifies to S32, drives swr_convert through ctypes, and checks
the md5 pins of the reference's decode test (reference tests/test_decode.c:35-36,55-56) before writing
the stereo fixture, whose force vector the reference pins in tests/test_analyze.c:59-78.
Run in the build container only (needs /root/reference and the opencv wheel): python tools/make_golden_s32.py
"""
import ctypes, glob, hashlib, os, subprocess, sys, tempfile
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBS = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs"
so = os.path.join(tempfile.mkdtemp(), "libflacrd.so")
subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(REPO, "bliss_b200/host/flac_reader.c")], check=True)
rd = ctypes.CDLL(so)
class PcmFile(ctypes.Structure):
    _fields_ = [("samples", ctypes.POINTER(ctypes.c_int32)), ("samples16", ctypes.POINTER(ctypes.c_int16)), ("n_frames", ctypes.c_size_t), ("channels", ctypes.c_int),
                ("sample_rate", ctypes.c_int), ("bits_per_sample", ctypes.c_int), ("is_float", ctypes.c_int), ("container", ctypes.c_int),
                ("file_bytes", ctypes.c_uint64), ("md5", ctypes.c_uint8 * 16)] + [(k, ctypes.c_char_p) for k in ("artist", "title", "album", "tracknumber", "genre")] + [
                    ("resampled16", ctypes.POINTER(ctypes.c_int16)), ("resampled_frames", ctypes.c_size_t)]
rd.blx_pcm_file_read.argtypes = [ctypes.c_char_p, ctypes.POINTER(PcmFile)]
def read(path):
    f = PcmFile()
    assert rd.blx_pcm_file_read(path.encode(), ctypes.byref(f)) == 0
    assert rd.blx_pcm_file_samples32(ctypes.byref(f)) == 0
    a = np.ctypeslib.as_array(f.samples, (f.n_frames * f.channels,)).copy()
    return a, f.n_frames, f.channels, f.sample_rate, f.bits_per_sample
for name in ("libdrm", "libcrypto", "libssl"):
    for p in glob.glob(os.path.join(LIBS, name + "*")):
        try: ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
        except OSError as e: print("preload", p, e)
avutil = ctypes.CDLL(glob.glob(os.path.join(LIBS, "libavutil-*"))[0], mode=ctypes.RTLD_GLOBAL)
swr = ctypes.CDLL(glob.glob(os.path.join(LIBS, "libswresample-*"))[0], mode=ctypes.RTLD_GLOBAL)
swr.swr_alloc.restype = ctypes.c_void_p
avutil.av_opt_set.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
avutil.av_opt_set_int.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int]
swr.swr_init.argtypes = [ctypes.c_void_p]
swr.swr_convert.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
def resample(x32, channels, rate_in):
    ctx = swr.swr_alloc()
    lay = b"stereo" if channels == 2 else b"mono"
    assert avutil.av_opt_set(ctx, b"in_chlayout", lay, 0) == 0
    assert avutil.av_opt_set(ctx, b"out_chlayout", b"stereo", 0) == 0
    for k, v in ((b"in_sample_rate", rate_in), (b"out_sample_rate", 22050), (b"in_sample_fmt", 2), (b"out_sample_fmt", 1)):
        assert avutil.av_opt_set_int(ctx, k, v, 0) == 0, k
    assert swr.swr_init(ctx) == 0
    n = len(x32) // channels
    out = []
    BLK = 4608
    obuf = np.zeros(2 * (BLK + 4096), dtype=np.int16)
    def conv(inp, n_in):
        op = (ctypes.c_void_p * 1)(obuf.ctypes.data)
        if inp is None:
            got = swr.swr_convert(ctx, op, BLK + 4096, None, 0)
        else:
            ip = (ctypes.c_void_p * 1)(inp.ctypes.data)
            got = swr.swr_convert(ctx, op, BLK + 4096, ip, n_in)
        assert got >= 0
        if got: out.append(obuf[:2 * got].copy())
        return got
    for s in range(0, n, BLK):
        blk = np.ascontiguousarray(x32[s * channels:(s + min(BLK, n - s)) * channels])
        conv(blk, len(blk) // channels)
    while conv(None, 0) > 0:
        pass
    return np.concatenate(out)
for fn, pin in (("song_s32.flac", "eb9f31a7b9ed022d66ff82b76e7c3c18"), ("song_s32_mono.flac", "747dbfcd75bebc23ebe2024935aede36")):
    a, n, ch, rate, bps = read("/root/reference/audio/" + fn)
    print(fn, n, ch, rate, bps)
    x32 = (a.astype(np.int64) << (32 - bps)).astype(np.int32)
    pcm = resample(x32, ch, rate)
    md5 = hashlib.md5(pcm.tobytes()).hexdigest()
    print(" ->", len(pcm), md5, "PIN OK" if md5 == pin else "pin mismatch " + pin)
    assert md5 == pin
    if fn == "song_s32.flac":
        np.savez_compressed(os.path.join(REPO, "tests", "golden", "song_s32_pcm.npz"), pcm=pcm)
