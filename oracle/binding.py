"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the CPU checkers.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product package (bliss_b200/).

  Oracle   -> oracle/liboracle.so        our C restatement (bliss_oracle.c, frontend.c)
  RefLib   -> oracle/_ref/libbliss_ref.so the reference's analyser sources compiled
              verbatim (built in the authoring container; travels prebuilt)
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libbliss_ref.so")

_i16p = ctypes.POINTER(ctypes.c_int16)
_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(quiet=True):
    """Compile liboracle.so (always) and _ref (only where /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class ForceVector(ctypes.Structure):
    _fields_ = [("tempo", ctypes.c_float), ("amplitude", ctypes.c_float),
                ("frequency", ctypes.c_float), ("attack", ctypes.c_float)]


class EnvelopeResult(ctypes.Structure):
    _fields_ = [("tempo", ctypes.c_float), ("attack", ctypes.c_float)]


class BlSong(ctypes.Structure):
    """struct bl_song, reference include/bliss.h:49-67 (120 bytes on LP64)."""
    _fields_ = [("force", ctypes.c_float), ("force_vector", ForceVector),
                ("sample_array", ctypes.c_void_p), ("channels", ctypes.c_int),
                ("nSamples", ctypes.c_int), ("sample_rate", ctypes.c_int),
                ("bitrate", ctypes.c_int), ("nb_bytes_per_sample", ctypes.c_int),
                ("calm_or_loud", ctypes.c_int), ("resampled", ctypes.c_int),
                ("duration", ctypes.c_uint64), ("filename", ctypes.c_void_p),
                ("artist", ctypes.c_void_p), ("title", ctypes.c_void_p),
                ("album", ctypes.c_void_p), ("tracknumber", ctypes.c_void_p),
                ("genre", ctypes.c_void_p)]


class OrcResult(ctypes.Structure):
    _fields_ = [("force", ctypes.c_float), ("tempo", ctypes.c_float), ("amplitude", ctypes.c_float),
                ("frequency", ctypes.c_float), ("attack", ctypes.c_float),
                ("calm_or_loud", ctypes.c_int), ("beat", ctypes.c_int)]


def _pcm(a):
    a = np.ascontiguousarray(a, dtype=np.int16)
    return a, a.ctypes.data_as(_i16p)


class Oracle:
    """Our C restatement (oracle/bliss_oracle.c, oracle/frontend.c)."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        L = self.lib = ctypes.CDLL(ORACLE_SO)
        L.orc_frequency.restype = ctypes.c_float
        L.orc_frequency.argtypes = [_i16p, ctypes.c_int, ctypes.c_int]
        L.orc_frequency_spectrum.restype = ctypes.c_int
        L.orc_frequency_spectrum.argtypes = [_i16p, ctypes.c_int, ctypes.c_int, _f32p]
        L.orc_frequency_from_spectrum.restype = ctypes.c_float
        L.orc_frequency_from_spectrum.argtypes = [_f32p]
        L.orc_amplitude.restype = ctypes.c_float
        L.orc_amplitude.argtypes = [_i16p, ctypes.c_int]
        L.orc_amplitude_from_histogram.restype = ctypes.c_float
        L.orc_amplitude_from_histogram.argtypes = [_f32p, ctypes.c_int, ctypes.c_int]
        L.orc_mean.restype = ctypes.c_int
        L.orc_mean.argtypes = [_i16p, ctypes.c_int]
        L.orc_variance.restype = ctypes.c_int
        L.orc_variance.argtypes = [_i16p, ctypes.c_int, ctypes.c_int]
        L.orc_envelope_energy.restype = ctypes.c_int
        L.orc_envelope_energy.argtypes = [_i16p, ctypes.c_int, _f64p]
        L.orc_envelope_tail.restype = None
        L.orc_envelope_tail.argtypes = [_f64p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                                        ctypes.POINTER(ctypes.c_int), _f64p, _f32p, _f32p, _f64p]
        L.orc_rectangular_filter.restype = None
        L.orc_rectangular_filter.argtypes = [_f64p, _f64p, ctypes.c_int, ctypes.c_int]
        L.orc_analyze.restype = None
        L.orc_analyze.argtypes = [_i16p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                                  ctypes.POINTER(OrcResult)]
        L.orc_rating.restype = ctypes.c_float
        L.orc_rating.argtypes = [ctypes.c_float] * 4 + [ctypes.POINTER(ctypes.c_int)]
        L.orc_distance.restype = ctypes.c_float
        L.orc_distance.argtypes = [_f32p, _f32p]
        L.orc_cosine_similarity.restype = ctypes.c_float
        L.orc_cosine_similarity.argtypes = [_f32p, _f32p]
        L.orc_distance_matrix.restype = None
        L.orc_distance_matrix.argtypes = [_f32p, ctypes.c_int, _f32p]
        L.orc_frontend_f32.restype = ctypes.c_int
        L.orc_frontend_f32.argtypes = [_f32p, ctypes.c_long, _i16p]
        L.orc_resample_to_s16.restype = ctypes.c_longlong
        L.orc_resample_to_s16.argtypes = [ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_longlong, ctypes.c_int, _i16p]

    # -- analysers ---------------------------------------------------
    def frequency(self, pcm, channels=2):
        a, p = _pcm(pcm)
        return float(self.lib.orc_frequency(p, len(a), channels))

    def frequency_spectrum(self, pcm, channels=2):
        a, p = _pcm(pcm)
        ps = np.zeros(257, dtype=np.float32)
        self.lib.orc_frequency_spectrum(p, len(a), channels, ps.ctypes.data_as(_f32p))
        return ps

    def frequency_from_spectrum(self, ps):
        ps = np.array(ps, dtype=np.float32, copy=True)
        return float(self.lib.orc_frequency_from_spectrum(ps.ctypes.data_as(_f32p)))

    def amplitude(self, pcm):
        a, p = _pcm(pcm)
        return float(self.lib.orc_amplitude(p, len(a)))

    def mean_variance(self, pcm):
        a, p = _pcm(pcm)
        m = self.lib.orc_mean(p, len(a))
        return m, self.lib.orc_variance(p, len(a), m)

    def envelope_energy(self, pcm):
        a, p = _pcm(pcm)
        nb = 2 * (len(a) // 512)
        E = np.zeros(max(nb, 1), dtype=np.float64)
        self.lib.orc_envelope_energy(p, len(a), E.ctypes.data_as(_f64p))
        return E[:nb]

    def envelope_tail(self, E, n_samples, duration, want_signal=False):
        E = np.ascontiguousarray(E, dtype=np.float64)
        beat = ctypes.c_int(0)
        atk = ctypes.c_double(0)
        tempo = ctypes.c_float(0)
        attack = ctypes.c_float(0)
        ss = np.zeros(2 * len(E), dtype=np.float64) if want_signal else None
        self.lib.orc_envelope_tail(E.ctypes.data_as(_f64p), len(E), n_samples, duration,
                                   ctypes.byref(beat), ctypes.byref(atk), ctypes.byref(tempo),
                                   ctypes.byref(attack),
                                   ss.ctypes.data_as(_f64p) if want_signal else None)
        out = dict(beat=beat.value, atk_sum=atk.value, tempo=tempo.value, attack=attack.value)
        if want_signal:
            out["signal"] = ss
        return out

    def rectangular_filter(self, out, inp, width=19):
        out = np.array(out, dtype=np.float64, copy=True)
        inp = np.ascontiguousarray(inp, dtype=np.float64)
        self.lib.orc_rectangular_filter(out.ctypes.data_as(_f64p), inp.ctypes.data_as(_f64p), len(inp), width)
        return out

    def analyze(self, pcm, duration, channels=2):
        a, p = _pcm(pcm)
        r = OrcResult()
        self.lib.orc_analyze(p, len(a), channels, int(duration), ctypes.byref(r))
        return dict(force=r.force, tempo=r.tempo, amplitude=r.amplitude, frequency=r.frequency,
                    attack=r.attack, calm_or_loud=r.calm_or_loud, beat=r.beat)

    def rating(self, tempo, amplitude, frequency, attack):
        c = ctypes.c_int(0)
        f = self.lib.orc_rating(tempo, amplitude, frequency, attack, ctypes.byref(c))
        return float(f), c.value

    # -- distances ---------------------------------------------------
    def distance(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        return float(self.lib.orc_distance(a.ctypes.data_as(_f32p), b.ctypes.data_as(_f32p)))

    def cosine_similarity(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        return float(self.lib.orc_cosine_similarity(a.ctypes.data_as(_f32p), b.ctypes.data_as(_f32p)))

    def distance_matrix(self, v):
        v = np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 4)
        out = np.zeros((len(v), len(v)), dtype=np.float32)
        self.lib.orc_distance_matrix(v.ctypes.data_as(_f32p), len(v), out.ctypes.data_as(_f32p))
        return out

    # -- front-end ---------------------------------------------------
    def frontend_f32(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros(2 * (len(x) // 2), dtype=np.int16)
        self.lib.orc_frontend_f32(x.ctypes.data_as(_f32p), len(x), out.ctypes.data_as(_i16p))
        return out


    # -- decode-stage resampler (include/blx_resample.h) -----------------
    RS_S16, RS_S32, RS_F32, RS_U8 = 0, 1, 2, 3

    def resample_to_s16(self, samples, kind, bits, channels, in_rate):
        """samples: the reader's int32 array (interleaved; float32 bits for RS_F32). Returns int16 stereo interleaved
        at 22 050 Hz as libswresample with default options produces it."""
        a = np.ascontiguousarray(samples, dtype=np.int32)
        n = len(a) // channels
        p = a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        cnt = self.lib.orc_resample_to_s16(p, kind, bits, channels, n, in_rate, None)
        if cnt < 0:
            raise ValueError("unsupported resampling request")
        out = np.zeros(2 * cnt, dtype=np.int16)
        self.lib.orc_resample_to_s16(p, kind, bits, channels, n, in_rate, out.ctypes.data_as(_i16p))
        return out


class RefLib:
    """The reference's own analyser sources, compiled verbatim (oracle/_ref)."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(
                REF_SO + " missing: it is built from /root/reference by `make -C oracle ref` "
                "in the authoring container and travels prebuilt")
        L = self.lib = ctypes.CDLL(REF_SO)
        L.bl_amplitude_sort.restype = ctypes.c_float
        L.bl_amplitude_sort.argtypes = [ctypes.POINTER(BlSong)]
        L.bl_frequency_sort.restype = ctypes.c_float
        L.bl_frequency_sort.argtypes = [ctypes.POINTER(BlSong)]
        L.bl_envelope_sort.restype = None
        L.bl_envelope_sort.argtypes = [ctypes.POINTER(BlSong), ctypes.POINTER(EnvelopeResult)]
        L.bl_analyze.restype = ctypes.c_int
        L.bl_analyze.argtypes = [ctypes.c_char_p, ctypes.POINTER(BlSong)]
        L.bl_free_song.restype = None
        L.bl_free_song.argtypes = [ctypes.POINTER(BlSong)]
        L.bl_distance.restype = ctypes.c_float
        L.bl_distance.argtypes = [ForceVector, ForceVector]
        L.bl_cosine_similarity.restype = ctypes.c_float
        L.bl_cosine_similarity.argtypes = [ForceVector, ForceVector]
        L.bl_mean.restype = ctypes.c_int
        L.bl_mean.argtypes = [_i16p, ctypes.c_int]
        L.bl_variance.restype = ctypes.c_int
        L.bl_variance.argtypes = [_i16p, ctypes.c_int, ctypes.c_int]
        L.bl_rectangular_filter.restype = None
        L.bl_rectangular_filter.argtypes = [_f64p, _f64p, ctypes.c_int, ctypes.c_int]
        L.oracle_ref_set_pcm.restype = None
        L.oracle_ref_set_pcm.argtypes = [_i16p, ctypes.c_int, ctypes.c_uint64, ctypes.c_int]
        L.oracle_ref_sizeof_bl_song.restype = ctypes.c_size_t

    @staticmethod
    def make_song(pcm, duration, channels=2):
        a = np.ascontiguousarray(pcm, dtype=np.int16)
        s = BlSong()
        s.sample_array = a.ctypes.data
        s.nSamples = len(a)
        s.channels = channels
        s.sample_rate = 22050
        s.nb_bytes_per_sample = 2
        s.duration = int(duration)
        return s, a  # keep `a` alive

    def analyze_pcm(self, pcm, duration, channels=2):
        """The three analysers + the rating of reference src/analyze.c:63-79, called the way
        bl_analyze does, on PCM already in memory (thread-safe: no decode stub involved)."""
        s, keep = self.make_song(pcm, duration, channels)
        amp = self.lib.bl_amplitude_sort(ctypes.byref(s))
        freq = self.lib.bl_frequency_sort(ctypes.byref(s))
        env = EnvelopeResult()
        self.lib.bl_envelope_sort(ctypes.byref(s), ctypes.byref(env))
        del keep
        return dict(tempo=env.tempo, amplitude=amp, frequency=freq, attack=env.attack)

    def bl_analyze(self, pcm, duration, channels=2):
        """The reference's bl_analyze end to end, with decode replaced by the in-memory stub."""
        a, p = _pcm(pcm)
        self.lib.oracle_ref_set_pcm(p, len(a), int(duration), channels)
        s = BlSong()
        rc = self.lib.bl_analyze(b"<memory>", ctypes.byref(s))
        out = dict(rc=rc, force=s.force, tempo=s.force_vector.tempo, amplitude=s.force_vector.amplitude,
                   frequency=s.force_vector.frequency, attack=s.force_vector.attack,
                   calm_or_loud=s.calm_or_loud, nSamples=s.nSamples, channels=s.channels)
        self.lib.bl_free_song(ctypes.byref(s))
        return out

    def distance(self, a, b):
        return float(self.lib.bl_distance(ForceVector(*map(float, a)), ForceVector(*map(float, b))))

    def cosine_similarity(self, a, b):
        return float(self.lib.bl_cosine_similarity(ForceVector(*map(float, a)), ForceVector(*map(float, b))))
