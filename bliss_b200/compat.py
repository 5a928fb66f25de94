"""The reference's Python package surface (`import bliss`) over the B200 library.

Mirrors reference python/bliss/{__init__,bl_song,distance,version}.py: the dict-like `bl_song` wrapper
of `struct bl_song`, `distance.distance` / `distance.cosine_similarity`, `version.version` and the
BL_* status codes, with ctypes instead of cffi (the reference compiles the C sources into its own
extension, reference python/build_bliss.py:21-33; here the in-tree libbliss.so is loaded). Usage:

    from bliss_b200 import compat as bliss
    with bliss.bl_song("song.flac") as song:
        print(song["force_vector"], song["title"])

Two binding bugs of the reference are not reproduced: `sample_array` is returned as the nSamples int16
values (reference bl_song.py:108-109 reads nSamples int8 values, half the buffer) and
`cosine_similarity` on two file names calls bl_cosine_similarity_file (reference bl_song.py:250 passes
file names to bl_cosine_similarity).

On top of the reference's surface, `playlist()` is the GPU form of the reference's playlist example
(reference python/examples/make_m3u_playlist.py:51-76: distances from a seed song to every song, then
argsort): the distances come from the library's row kernel, bit-identical to bl_distance.
"""
import ctypes
from collections.abc import Mapping

import numpy as np

from . import _lib as L

BL_LOUD, BL_CALM, BL_UNKNOWN, BL_UNEXPECTED, BL_OK = 0, 1, 2, -2, 0  # reference include/bliss.h:20-24

_STRINGS = ("filename", "artist", "title", "album", "tracknumber", "genre")


class bl_song(Mapping):
    """Dict-like wrapper of `struct bl_song` (reference python/bliss/bl_song.py:9-209)."""

    def __init__(self, filename=None, initializer=None, c_struct=None):
        self._lib = L.load()
        if c_struct is not None:
            self._c_struct = c_struct
        else:
            self._c_struct = L.BlSong()
            self._lib.bl_initialize_song(ctypes.byref(self._c_struct))
            if isinstance(initializer, dict):
                for k, v in initializer.items():
                    self.set(k, v)
        self._fields = [f[0] for f in L.BlSong._fields_]
        self._keepalive = {}
        if filename is not None:
            self.analyze(filename)

    # ---- Mapping interface
    def __getitem__(self, key):
        return self.get(key)

    def __setitem__(self, key, value):
        return self.set(key, value)

    def __len__(self):
        return len(self._fields)

    def __iter__(self):
        return iter(self._fields)

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.free()

    def __repr__(self):
        return {k: (self.get(k) if k != "sample_array" else "<%d samples>" % self._c_struct.nSamples)
                for k in self._fields}.__repr__()

    def get(self, key):
        if key not in self._fields:
            raise KeyError(key)
        value = getattr(self._c_struct, key)
        if key in _STRINGS:
            return value.decode("utf-8", "replace") if value is not None else None
        if key == "force_vector":
            return {"tempo": value.tempo, "amplitude": value.amplitude, "frequency": value.frequency,
                    "attack": value.attack}
        if key == "sample_array":
            n = self._c_struct.nSamples
            if not value or n <= 0:
                return []
            return np.ctypeslib.as_array(ctypes.cast(value, ctypes.POINTER(ctypes.c_int16)), (n,)).tolist()
        return value

    def set(self, key, value):
        if key not in self._fields:
            raise KeyError(key)
        if key in _STRINGS:
            # kept alive here, not on the C heap: free() detaches it before bl_free_song runs
            buf = ctypes.create_string_buffer(value.encode("utf-8")) if value is not None else None
            self._keepalive[key] = buf
            setattr(self._c_struct, key, ctypes.cast(buf, ctypes.c_char_p) if buf is not None else None)
        elif key == "force_vector":
            if value is None:
                return
            fv = self._c_struct.force_vector
            if isinstance(value, dict):
                value = (value["tempo"], value["amplitude"], value["frequency"], value["attack"])
            fv.tempo, fv.amplitude, fv.frequency, fv.attack = (float(x) for x in value)
        elif key == "sample_array":
            arr = np.ascontiguousarray(value, dtype=np.int16) if value is not None else None
            self._keepalive[key] = arr
            self._c_struct.sample_array = arr.ctypes.data if arr is not None else None
        else:
            setattr(self._c_struct, key, value)

    # ---- the library calls
    def decode(self, filename):
        """bl_audio_decode: load a file (FLAC / WAV here), no analysis."""
        return self._lib.bl_audio_decode(filename.encode("utf-8"), ctypes.byref(self._c_struct))

    def analyze(self, filename):
        """bl_analyze: decode + the three analysers on the GPU; returns BL_LOUD / BL_CALM / BL_UNKNOWN / BL_UNEXPECTED."""
        return self._lib.bl_analyze(filename.encode("utf-8"), ctypes.byref(self._c_struct))

    def envelope_analysis(self):
        result = L.EnvelopeResult()
        self._lib.bl_envelope_sort(ctypes.byref(self._c_struct), ctypes.byref(result))
        return {"tempo": result.tempo, "attack": result.attack}

    def amplitude_analysis(self):
        return self._lib.bl_amplitude_sort(ctypes.byref(self._c_struct))

    def frequency_analysis(self):
        return self._lib.bl_frequency_sort(ctypes.byref(self._c_struct))

    def free(self):
        """Release what the library allocated (bl_free_song); Python-owned members are detached first."""
        for k in list(self._keepalive):
            del self._keepalive[k]
            setattr(self._c_struct, k, None)
        self._lib.bl_free_song(ctypes.byref(self._c_struct))


def _fv(song):
    v = song["force_vector"]
    return L.ForceVector(v["tempo"], v["amplitude"], v["frequency"], v["attack"])


def distance(song1, song2):
    """reference python/bliss/distance.py:5-41: two file names (bl_distance_file) or two bl_song objects (bl_distance)."""
    lib = L.load()
    if isinstance(song1, str) and isinstance(song2, str):
        s1, s2 = L.BlSong(), L.BlSong()
        lib.bl_initialize_song(ctypes.byref(s1))
        lib.bl_initialize_song(ctypes.byref(s2))
        d = lib.bl_distance_file(song1.encode("utf-8"), song2.encode("utf-8"), ctypes.byref(s1), ctypes.byref(s2))
        return {"distance": d, "song1": bl_song(c_struct=s1), "song2": bl_song(c_struct=s2)}
    if isinstance(song1, bl_song) and isinstance(song2, bl_song):
        return {"distance": lib.bl_distance(_fv(song1), _fv(song2)), "song1": song1, "song2": song2}
    return {"distance": None, "song1": None, "song2": None}


def cosine_similarity(song1, song2):
    """reference python/bliss/distance.py:44-76."""
    lib = L.load()
    if isinstance(song1, str) and isinstance(song2, str):
        s1, s2 = L.BlSong(), L.BlSong()
        lib.bl_initialize_song(ctypes.byref(s1))
        lib.bl_initialize_song(ctypes.byref(s2))
        c = lib.bl_cosine_similarity_file(song1.encode("utf-8"), song2.encode("utf-8"), ctypes.byref(s1), ctypes.byref(s2))
        return {"similarity": c, "song1": bl_song(c_struct=s1), "song2": bl_song(c_struct=s2)}
    if isinstance(song1, bl_song) and isinstance(song2, bl_song):
        return {"similarity": lib.bl_cosine_similarity(_fv(song1), _fv(song2)), "song1": song1, "song2": song2}
    return {"similarity": None, "song1": None, "song2": None}


def version():
    """reference python/bliss/version.py:4-8."""
    return L.load().bl_version()


def playlist(engine, vectors, seed_index, k=None, metric="euclidean"):
    """Songs ordered by bl_distance to song `seed_index` (the seed itself first, distance 0), or, with
    metric="cosine", by decreasing bl_cosine_similarity (reference src/analyze.c:127-145).

    `vectors`: (n, 4) float32 [tempo, amplitude, frequency, attack]. The n values are one row of the
    all-pairs kernel (bit-identical to the reference's scalar functions); ties keep index order. Returns
    (indices, values) as numpy arrays, cut to the first k if given."""
    import torch
    v = torch.as_tensor(np.ascontiguousarray(vectors, dtype=np.float32).reshape(-1, 4)).cuda()
    n = v.shape[0]
    if not 0 <= seed_index < n:
        raise IndexError(seed_index)
    row = torch.empty(n, dtype=torch.float32, device=v.device)
    st = torch.cuda.current_stream(v.device).cuda_stream
    cosine = metric == "cosine"
    if not cosine and metric != "euclidean":
        raise ValueError(metric)
    engine.distance_rows_device(v.data_ptr(), n, int(seed_index), 1, row.data_ptr(), cosine=cosine, stream=st)
    order = torch.argsort(row, stable=True, descending=cosine)
    if k is not None:
        order = order[:k]
    return order.cpu().numpy(), row[order].cpu().numpy()
