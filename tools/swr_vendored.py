"""ctypes driver for the libswresample that opencv-python-headless vendors (build container only; FFmpeg itself is not
in the image). Used by tools/make_golden_resample.py and tools/make_golden_s32.py-style scripts to pin the decode-stage
resampler specification (include/blx_resample.h) against the real library. Drives swr_convert the way the reference
does (reference src/decode.c:388-392): blocks of one decoded frame, then flush until empty."""
import ctypes, glob, os
import numpy as np
LIBS = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs"
for name in ("libdrm", "libcrypto", "libssl"):
    for p in glob.glob(os.path.join(LIBS, name + "*")):
        try: ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
        except OSError as e: pass
avutil = ctypes.CDLL(glob.glob(os.path.join(LIBS, "libavutil-*"))[0], mode=ctypes.RTLD_GLOBAL)
swr = ctypes.CDLL(glob.glob(os.path.join(LIBS, "libswresample-*"))[0], mode=ctypes.RTLD_GLOBAL)
swr.swr_alloc.restype = ctypes.c_void_p
avutil.av_opt_set.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
avutil.av_opt_set_int.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int]
swr.swr_init.argtypes = [ctypes.c_void_p]
swr.swr_free.argtypes = [ctypes.POINTER(ctypes.c_void_p)]
swr.swr_convert.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
avutil.av_get_cpu_flags.restype = ctypes.c_int
FMT = {"u8":0,"s16":1,"s32":2,"flt":3,"dbl":4}
DT = {"u8":np.uint8,"s16":np.int16,"s32":np.int32,"flt":np.float32,"dbl":np.float64}
def resample(x, channels, rate_in, in_fmt="s32", out_fmt="s16", rate_out=22050, out_layout=b"stereo", blk=4608, opts=()):
    ctx = swr.swr_alloc()
    lay = b"stereo" if channels == 2 else b"mono"
    assert avutil.av_opt_set(ctx, b"in_chlayout", lay, 0) == 0
    assert avutil.av_opt_set(ctx, b"out_chlayout", out_layout, 0) == 0
    for k, v in ((b"in_sample_rate", rate_in), (b"out_sample_rate", rate_out), (b"in_sample_fmt", FMT[in_fmt]), (b"out_sample_fmt", FMT[out_fmt])) + tuple(opts):
        assert avutil.av_opt_set_int(ctx, k, v, 0) == 0, k
    assert swr.swr_init(ctx) == 0
    och = 2 if out_layout == b"stereo" else 1
    x = np.ascontiguousarray(x, dtype=DT[in_fmt])
    n = len(x) // channels
    out = []
    cap = int(blk * rate_out / rate_in) + 8192
    obuf = np.zeros(och * cap, dtype=DT[out_fmt])
    def conv(inp, n_in):
        op = (ctypes.c_void_p * 1)(obuf.ctypes.data)
        if inp is None: got = swr.swr_convert(ctx, op, cap, None, 0)
        else:
            ip = (ctypes.c_void_p * 1)(inp.ctypes.data)
            got = swr.swr_convert(ctx, op, cap, ip, n_in)
        assert got >= 0
        if got: out.append(obuf[:och * got].copy())
        return got
    for s in range(0, n, blk):
        b = np.ascontiguousarray(x[s * channels:(s + min(blk, n - s)) * channels])
        conv(b, len(b) // channels)
    while conv(None, 0) > 0: pass
    c = ctypes.c_void_p(ctx); swr.swr_free(ctypes.byref(c))
    return np.concatenate(out) if out else np.zeros(0, DT[out_fmt])
