"""CPU: the drop-in library loads and exports every symbol include/bliss.h and include/blx.h declare,
struct layouts match the reference ABI, and the product fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

import bliss_b200
from bliss_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = bliss_b200.load()
    bl = _declared("bliss.h", "bl_")
    blx = _declared("blx.h", "blx_")
    assert len(bl) == 15, bl  # the 15 prototypes of reference include/bliss.h:80-290
    assert sorted(_lib.BLISS_H_SYMBOLS) == bl
    assert sorted(_lib.BLX_H_SYMBOLS) == blx
    for name in bl + blx:
        assert hasattr(L, name), name


def test_struct_abi_matches_reference():
    # 120-byte bl_song with these offsets; 16-byte force_vector_s
    S = bliss_b200.BlSong
    assert ctypes.sizeof(S) == 120 and ctypes.sizeof(bliss_b200.ForceVector) == 16
    offs = {n: getattr(S, n).offset for n, _ in S._fields_}
    assert offs == dict(force=0, force_vector=4, sample_array=24, channels=32, nSamples=36, sample_rate=40,
                        bitrate=44, nb_bytes_per_sample=48, calm_or_loud=52, resampled=56, duration=64,
                        filename=72, artist=80, title=88, album=96, tracknumber=104, genre=112)
    assert ctypes.sizeof(bliss_b200.BlxResult) == 32


def test_struct_abi_matches_reference_build(reflib):
    assert reflib.lib.oracle_ref_sizeof_bl_song() == ctypes.sizeof(bliss_b200.BlSong)


def test_scalar_distance_host_api(oracle):
    L = bliss_b200.load()
    a = bliss_b200.ForceVector(1.5, -2.25, 3.125, 0.1)
    b = bliss_b200.ForceVector(-0.5, 7.0, 3.0, -9.9)
    av, bv = [1.5, -2.25, 3.125, 0.1], [-0.5, 7.0, 3.0, -9.9]
    assert L.bl_distance(a, b) == oracle.distance(av, bv)
    assert L.bl_cosine_similarity(a, b) == oracle.cosine_similarity(av, bv)


def test_lifecycle_helpers():
    L = bliss_b200.load()
    s = bliss_b200.BlSong()
    L.bl_initialize_song(ctypes.byref(s))
    assert s.sample_array is None and s.artist is None
    L.bl_free_song(ctypes.byref(s))  # freeing an initialised, empty song is legal (reference examples/analyze.c:15-17,50-52)
    assert abs(L.bl_version() - 1.2) < 1e-6


def test_decode_errors():
    L = bliss_b200.load()
    s = bliss_b200.BlSong()
    assert L.bl_audio_decode(b"/nonexistent/file.flac", ctypes.byref(s)) == -2  # BL_UNEXPECTED
    assert L.bl_analyze(b"/nonexistent/file.flac", ctypes.byref(s)) == -2


def test_no_cpu_fallback_without_gpu():
    L = bliss_b200.load()
    if L.blx_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(bliss_b200.BlxError):
        bliss_b200.Engine(0)
    # the bliss.h analysers report failure instead of computing on the CPU
    import numpy as np
    pcm = np.ones(50000, dtype=np.int16)
    s = bliss_b200.BlSong()
    s.sample_array = pcm.ctypes.data
    s.nSamples = len(pcm)
    s.channels = 2
    s.duration = 1
    import math
    assert math.isnan(L.bl_amplitude_sort(ctypes.byref(s)))
