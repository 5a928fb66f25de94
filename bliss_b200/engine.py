"""Python face of the device C-ABI (include/blx.h): one Engine per GPU / process.

Host-buffer calls take numpy arrays; device-resident calls take raw device pointers
(e.g. torch.Tensor.data_ptr()) — PyTorch is only the allocator / stream / collective plumbing.
"""
import ctypes

import numpy as np

from . import _lib as L

DO_AMPLITUDE, DO_FREQUENCY, DO_ENVELOPE, DO_ALL = 0x1, 0x2, 0x4, 0x7
FMT_S16, FMT_F32 = 0, 1
ALIGN_ELEMS = 64
K_PASS1, K_EPILOGUE, K_ENVELOPE, K_TAIL, K_DISTANCE, K_COUNT = 0, 1, 2, 3, 4, 5
SONG_TOO_SHORT, SONG_SILENT, SONG_FLAT = 0x1, 0x2, 0x4
DEBUG_SLOW_CHAIN = 0x1

RESULT_DTYPE = np.dtype([("tempo", "<f4"), ("amplitude", "<f4"), ("frequency", "<f4"), ("attack", "<f4"),
                         ("force", "<f4"), ("calm_or_loud", "<i4"), ("beat", "<i4"), ("status", "<i4")])
assert RESULT_DTYPE.itemsize == ctypes.sizeof(L.BlxResult) == 32


class BlxError(RuntimeError):
    pass


def _c_array(ctype, values):
    """A ctypes array of `values`; an array of the right type passes through (callers that reuse the same offsets /
    lengths for many launches build it once)."""
    if isinstance(values, ctypes.Array) and values._type_ is ctype:
        return values
    return (ctype * len(values))(*[int(v) for v in values])


def _stream_arg(stream):
    """cudaStream_t for the C-ABI. None = the engine's own stream. A handle of 0 is CUDA's legacy default stream (what
    torch.cuda.current_stream().cuda_stream returns outside a stream context): it is passed as cudaStreamLegacy (0x1),
    because NULL already means "the engine's own stream" - the work must be ordered with the caller's stream, not beside it."""
    if stream is None:
        return None
    return ctypes.c_void_p(int(stream) or 1)


class Engine:
    def __init__(self, device=0, chunk_bytes=None):
        self._lib = L.load()
        h = ctypes.c_void_p()
        rc = self._lib.blx_init(int(device), ctypes.byref(h))
        if rc != 0:
            raise BlxError(f"blx_init(device={device}) failed ({rc}): {self._lib.blx_last_error().decode()}")
        self._h = h
        self.device = int(device)
        if chunk_bytes:
            self._ck(self._lib.blx_configure(self._h, int(chunk_bytes)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.blx_shutdown(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise BlxError(f"blx call failed ({rc}): {self._lib.blx_last_error().decode()}")

    # ------------------------------------------------------------------ host buffers
    def analyze_s16(self, songs, durations, channels=None, what=DO_ALL):
        """songs: list of int16 numpy arrays (interleaved). Returns a RESULT_DTYPE array."""
        n = len(songs)
        songs = [np.ascontiguousarray(s, dtype=np.int16) for s in songs]
        ptrs = (ctypes.c_void_p * n)(*[s.ctypes.data for s in songs])
        lens = (ctypes.c_int * n)(*[len(s) for s in songs])
        durs = (ctypes.c_uint64 * n)(*[int(d) for d in durations])
        chs = (ctypes.c_int * n)(*[int(c) for c in channels]) if channels is not None else None
        out = np.zeros(n, dtype=RESULT_DTYPE)
        self._ck(self._lib.blx_analyze_batch_s16(self._h, ptrs, lens, chs, durs, n, what,
                                                 out.ctypes.data_as(ctypes.POINTER(L.BlxResult))))
        return out

    def analyze_f32(self, songs, what=DO_ALL):
        """songs: list of float32 numpy arrays (44.1 kHz mono)."""
        n = len(songs)
        songs = [np.ascontiguousarray(s, dtype=np.float32) for s in songs]
        ptrs = (ctypes.c_void_p * n)(*[s.ctypes.data for s in songs])
        lens = (ctypes.c_int64 * n)(*[len(s) for s in songs])
        out = np.zeros(n, dtype=RESULT_DTYPE)
        self._ck(self._lib.blx_analyze_batch_f32(self._h, ptrs, lens, n, what,
                                                 out.ctypes.data_as(ctypes.POINTER(L.BlxResult))))
        return out

    def analyze_f32_exact(self, songs, in_rate=44100, channels=1, what=DO_ALL):
        """float32 songs through the decode-stage resampler (libswresample-exact) + the native int16 analysis:
        what bl_analyze gives for a float file of that rate."""
        n = len(songs)
        songs = [np.ascontiguousarray(s, dtype=np.float32) for s in songs]
        ptrs = (ctypes.c_void_p * n)(*[s.ctypes.data for s in songs])
        lens = (ctypes.c_int64 * n)(*[len(s) // channels for s in songs])
        out = np.zeros(n, dtype=RESULT_DTYPE)
        self._ck(self._lib.blx_analyze_batch_f32_exact(self._h, ptrs, lens, int(channels), int(in_rate), n, what,
                                                       out.ctypes.data_as(ctypes.POINTER(L.BlxResult))))
        return out

    def analyze_host_ptrs(self, fmt, ptrs, lens, durations=None, channels=None, what=DO_ALL, out=None):
        """Like analyze_s16 / analyze_f32 on raw host addresses (e.g. pinned torch tensors)."""
        n = len(ptrs)
        cptrs = (ctypes.c_void_p * n)(*[int(p) for p in ptrs])
        if out is None:
            out = np.zeros(n, dtype=RESULT_DTYPE)
        res = out.ctypes.data_as(ctypes.POINTER(L.BlxResult))
        if fmt == FMT_F32:
            clens = (ctypes.c_int64 * n)(*[int(x) for x in lens])
            self._ck(self._lib.blx_analyze_batch_f32(self._h, cptrs, clens, n, what, res))
        else:
            clens = (ctypes.c_int * n)(*[int(x) for x in lens])
            durs = (ctypes.c_uint64 * n)(*[int(d) for d in durations])
            chs = (ctypes.c_int * n)(*[int(c) for c in channels]) if channels is not None else None
            self._ck(self._lib.blx_analyze_batch_s16(self._h, cptrs, clens, chs, durs, n, what, res))
        return out

    # ------------------------------------------------------------------ device resident
    def analyze_device(self, fmt, d_pcm, offsets, lengths, d_out, durations=None, channels=None, what=DO_ALL,
                       stream=None, wait=True):
        """wait=False: blx_analyze_device_async - `stream` is not made to wait for the results; call join(stream)
        after the last batch of the job."""
        n = len(offsets)
        offs = _c_array(ctypes.c_int64, offsets)
        lens = _c_array(ctypes.c_int64, lengths)
        durs = _c_array(ctypes.c_uint64, durations) if durations is not None else None
        chs = _c_array(ctypes.c_int, channels) if channels is not None else None
        f = self._lib.blx_analyze_device if wait else self._lib.blx_analyze_device_async
        self._ck(f(self._h, fmt, ctypes.c_void_p(int(d_pcm)), offs, lens, chs, durs, n, what, ctypes.c_void_p(int(d_out)),
                   _stream_arg(stream)))

    def join(self, stream=None):
        """Makes `stream` wait for every analysis enqueued so far (after analyze_device(..., wait=False))."""
        self._ck(self._lib.blx_join(self._h, _stream_arg(stream)))

    def spectral_device(self, fmt, d_pcm, offsets, lengths, d_frequency, channels=None, stream=None):
        n = len(offsets)
        offs = _c_array(ctypes.c_int64, offsets)
        lens = _c_array(ctypes.c_int64, lengths)
        chs = _c_array(ctypes.c_int, channels) if channels is not None else None
        self._ck(self._lib.blx_spectral_device(self._h, fmt, ctypes.c_void_p(int(d_pcm)), offs, lens, chs, n,
                                               ctypes.c_void_p(int(d_frequency)),
                                               _stream_arg(stream)))

    # ------------------------------------------------------------------ distances
    def distance_matrix(self, vectors, cosine=False):
        v = np.ascontiguousarray(vectors, dtype=np.float32).reshape(-1, 4)
        out = np.zeros((len(v), len(v)), dtype=np.float32)
        f = self._lib.blx_cosine_matrix if cosine else self._lib.blx_distance_matrix
        self._ck(f(self._h, v.ctypes.data_as(L.c_f32p), len(v), out.ctypes.data_as(L.c_f32p)))
        return out

    def distance_rows_device(self, d_vectors, n, row0, n_rows, d_out, cosine=False, stream=None):
        self._ck(self._lib.blx_distance_rows_device(self._h, ctypes.c_void_p(int(d_vectors)), n, row0, n_rows,
                                                    1 if cosine else 0, ctypes.c_void_p(int(d_out)),
                                                    _stream_arg(stream)))

    def distance_nearest_device(self, d_vectors, n, row0, n_rows, d_index, d_dist, d_sum, stream=None):
        self._ck(self._lib.blx_distance_nearest_device(
            self._h, ctypes.c_void_p(int(d_vectors)), n, row0, n_rows,
            ctypes.c_void_p(int(d_index)) if d_index else None, ctypes.c_void_p(int(d_dist)) if d_dist else None,
            ctypes.c_void_p(int(d_sum)) if d_sum else None, _stream_arg(stream)))

    def cosine_nearest_device(self, d_vectors, n, row0, n_rows, d_index, d_similarity, stream=None):
        self._ck(self._lib.blx_cosine_nearest_device(
            self._h, ctypes.c_void_p(int(d_vectors)), n, row0, n_rows, ctypes.c_void_p(int(d_index)) if d_index else None,
            ctypes.c_void_p(int(d_similarity)) if d_similarity else None, _stream_arg(stream)))

    # ------------------------------------------------------------------ helpers
    def mean_variance(self, pcm, mean_in=None):
        a = np.ascontiguousarray(pcm, dtype=np.int16)
        m, v = ctypes.c_int(0), ctypes.c_int(0)
        mi = ctypes.byref(ctypes.c_int(int(mean_in))) if mean_in is not None else None
        self._ck(self._lib.blx_mean_variance_s16(self._h, a.ctypes.data_as(L.c_i16p), len(a), mi,
                                                 ctypes.byref(m), ctypes.byref(v)))
        return m.value, v.value

    def rectangular_filter(self, out, inp, width=19):
        out = np.array(out, dtype=np.float64, copy=True)
        inp = np.ascontiguousarray(inp, dtype=np.float64)
        self._ck(self._lib.blx_rectangular_filter(self._h, out.ctypes.data_as(L.c_f64p), inp.ctypes.data_as(L.c_f64p),
                                                  len(inp), width))
        return out

    def frontend_f32(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros(2 * (len(x) // 2), dtype=np.int16)
        self._ck(self._lib.blx_frontend_f32(self._h, x.ctypes.data_as(L.c_f32p), len(x), out.ctypes.data_as(L.c_i16p)))
        return out

    RS_S16, RS_S32, RS_F32, RS_U8 = 0, 1, 2, 3

    def resample_to_s16(self, samples, kind, bits, channels, in_rate):
        """Decode-stage resampler (include/blx_resample.h): the reader's int32 samples -> int16 / 22 050 Hz / stereo."""
        a = np.ascontiguousarray(samples, dtype=np.int32)
        n = len(a) // channels
        p = a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        cnt = ctypes.c_int64(0)
        self._ck(self._lib.blx_resample_to_s16(self._h, p, kind, bits, channels, n, in_rate, None, 0, ctypes.byref(cnt)))
        out = np.zeros(2 * cnt.value, dtype=np.int16)
        if cnt.value:
            self._ck(self._lib.blx_resample_to_s16(self._h, p, kind, bits, channels, n, in_rate, out.ctypes.data_as(L.c_i16p),
                                                   cnt.value, ctypes.byref(cnt)))
        return out

    def resample_s16_to_s16(self, samples, channels, in_rate):
        """The same for 16-bit sources stored as int16 (what bl_audio_decode uses for 16-bit WAVE files)."""
        a = np.ascontiguousarray(samples, dtype=np.int16)
        n = len(a) // channels
        p = a.ctypes.data_as(L.c_i16p)
        cnt = ctypes.c_int64(0)
        self._ck(self._lib.blx_resample_s16_to_s16(self._h, p, channels, n, in_rate, None, 0, ctypes.byref(cnt)))
        out = np.zeros(2 * cnt.value, dtype=np.int16)
        if cnt.value:
            self._ck(self._lib.blx_resample_s16_to_s16(self._h, p, channels, n, in_rate, out.ctypes.data_as(L.c_i16p), cnt.value,
                                                       ctypes.byref(cnt)))
        return out

    def envelope_energy(self, pcm):
        a = np.ascontiguousarray(pcm, dtype=np.int16)
        nb = 2 * (len(a) // 512)
        E = np.zeros(max(nb, 1), dtype=np.float64)
        self._ck(self._lib.blx_envelope_energy_s16(self._h, a.ctypes.data_as(L.c_i16p), len(a), E.ctypes.data_as(L.c_f64p)))
        return E[:nb]

    def envelope_energy_f32(self, x):
        """Hop energies of a 44.1 kHz mono float32 song through the fused path (front-end + doubled-mono envelope)."""
        a = np.ascontiguousarray(x, dtype=np.float32)
        nb = 2 * ((2 * (len(a) // 2)) // 512)
        E = np.zeros(max(nb, 1), dtype=np.float64)
        self._ck(self._lib.blx_envelope_energy_f32(self._h, a.ctypes.data_as(L.c_f32p), len(a), E.ctypes.data_as(L.c_f64p)))
        return E[:nb]

    def frequency_spectrum(self, pcm, channels=2):
        """Per-bin power (sum over frames of |X_d|^2, d = 1..255) before the scalar epilogue; 257 floats."""
        a = np.ascontiguousarray(pcm, dtype=np.int16)
        ps = np.zeros(257, dtype=np.float32)
        self._ck(self._lib.blx_frequency_spectrum_s16(self._h, a.ctypes.data_as(L.c_i16p), len(a), int(channels),
                                                      ps.ctypes.data_as(L.c_f32p)))
        return ps

    def histogram(self, pcm):
        """(counts of sample values -1904..+1902, first non-zero index, last non-zero index)."""
        a = np.ascontiguousarray(pcm, dtype=np.int16)
        h = np.zeros(3807, dtype=np.uint32)
        first, last = ctypes.c_int(0), ctypes.c_int(0)
        self._ck(self._lib.blx_histogram_s16(self._h, a.ctypes.data_as(L.c_i16p), len(a),
                                             h.ctypes.data_as(ctypes.POINTER(ctypes.c_uint)), ctypes.byref(first),
                                             ctypes.byref(last)))
        return h, first.value, last.value

    def envelope_tail(self, energy, n_samples, duration):
        """The sequential tail alone on given hop energies: dict(beat, tempo, attack)."""
        E = np.ascontiguousarray(energy, dtype=np.float64)
        beat, tempo, attack = ctypes.c_int(0), ctypes.c_float(0), ctypes.c_float(0)
        self._ck(self._lib.blx_envelope_tail(self._h, E.ctypes.data_as(L.c_f64p), len(E), int(n_samples), int(duration),
                                             ctypes.byref(beat), ctypes.byref(tempo), ctypes.byref(attack)))
        return dict(beat=beat.value, tempo=tempo.value, attack=attack.value)

    def configure_sub_batch(self, songs):
        """Songs per kernel sequence of the device-resident entry points (results do not depend on it)."""
        self._ck(self._lib.blx_configure_sub_batch(self._h, int(songs)))

    def debug_flags(self, flags):
        """Test hooks (include/blx.h BLX_DEBUG_*); results must not depend on them."""
        self._ck(self._lib.blx_debug_flags(self._h, int(flags)))

    # ------------------------------------------------------------------ measurement
    def profile(self, on=True):
        self._ck(self._lib.blx_profile_enable(self._h, 1 if on else 0))

    def profile_reset(self):
        self._ck(self._lib.blx_profile_reset(self._h))

    def profile_read(self):
        ms = (ctypes.c_float * K_COUNT)()
        n = (ctypes.c_int * K_COUNT)()
        self._ck(self._lib.blx_profile_read(self._h, ms, n))
        return {self._lib.blx_kernel_name(i).decode(): (float(ms[i]), int(n[i])) for i in range(K_COUNT)}

    def measure_fp64_peak(self):
        """Measured DFMA throughput of this device right now, TFLOP/s (the envelope kernel's roofline denominator)."""
        tf, mhz = ctypes.c_double(0), ctypes.c_double(0)
        self._ck(self._lib.blx_measure_fp64_peak(self._h, ctypes.byref(tf), ctypes.byref(mhz)))
        return tf.value

    def launch_count(self):
        return int(self._lib.blx_launch_count(self._h))
