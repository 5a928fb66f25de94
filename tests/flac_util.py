"""Host-side file reader of the product (bliss_b200/host/flac_reader.c: FLAC + RIFF/WAVE, plain C, no GPU) for tests
and fixture generators."""
import ctypes

import numpy as np

import bliss_b200


class PcmFile(ctypes.Structure):
    _fields_ = [("samples", ctypes.POINTER(ctypes.c_int32)), ("samples16", ctypes.POINTER(ctypes.c_int16)), ("n_frames", ctypes.c_size_t), ("channels", ctypes.c_int),
                ("sample_rate", ctypes.c_int), ("bits_per_sample", ctypes.c_int), ("is_float", ctypes.c_int),
                ("container", ctypes.c_int), ("file_bytes", ctypes.c_uint64), ("md5", ctypes.c_uint8 * 16)] + [
                    (k, ctypes.c_char_p) for k in ("artist", "title", "album", "tracknumber", "genre")]


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
