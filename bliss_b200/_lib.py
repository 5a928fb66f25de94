"""ctypes loader for the in-tree libbliss.so (bliss.h API + blx.h device C-ABI).

Fails loudly: there is no Python / CPU implementation to fall back to.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BLISS_B200_LIB: another build of the same library (kernel A/B experiments, tools/ab_envelope.sh)
LIB_PATH = os.environ.get("BLISS_B200_LIB") or os.path.join(_HERE, "libbliss.so")

c_i16p = ctypes.POINTER(ctypes.c_int16)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_i32p = ctypes.POINTER(ctypes.c_int)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u64p = ctypes.POINTER(ctypes.c_uint64)


class ForceVector(ctypes.Structure):
    """struct force_vector_s (reference include/bliss.h:26-31)."""
    _fields_ = [("tempo", ctypes.c_float), ("amplitude", ctypes.c_float),
                ("frequency", ctypes.c_float), ("attack", ctypes.c_float)]


class EnvelopeResult(ctypes.Structure):
    """struct envelope_result_s (reference include/bliss.h:34-37)."""
    _fields_ = [("tempo", ctypes.c_float), ("attack", ctypes.c_float)]


class BlSong(ctypes.Structure):
    """struct bl_song (reference include/bliss.h:49-67), 120 bytes on LP64."""
    _fields_ = [("force", ctypes.c_float), ("force_vector", ForceVector),
                ("sample_array", ctypes.c_void_p), ("channels", ctypes.c_int),
                ("nSamples", ctypes.c_int), ("sample_rate", ctypes.c_int),
                ("bitrate", ctypes.c_int), ("nb_bytes_per_sample", ctypes.c_int),
                ("calm_or_loud", ctypes.c_int), ("resampled", ctypes.c_int),
                ("duration", ctypes.c_uint64), ("filename", ctypes.c_char_p),
                ("artist", ctypes.c_char_p), ("title", ctypes.c_char_p),
                ("album", ctypes.c_char_p), ("tracknumber", ctypes.c_char_p),
                ("genre", ctypes.c_char_p)]


class BlxResult(ctypes.Structure):
    """blx_result (include/blx.h), 32 bytes."""
    _fields_ = [("tempo", ctypes.c_float), ("amplitude", ctypes.c_float),
                ("frequency", ctypes.c_float), ("attack", ctypes.c_float),
                ("force", ctypes.c_float), ("calm_or_loud", ctypes.c_int),
                ("beat", ctypes.c_int), ("status", ctypes.c_int)]


# every symbol include/bliss.h and include/blx.h declare
BLISS_H_SYMBOLS = [
    "bl_analyze", "bl_distance_file", "bl_distance", "bl_cosine_similarity_file", "bl_cosine_similarity",
    "bl_envelope_sort", "bl_amplitude_sort", "bl_frequency_sort", "bl_audio_decode", "bl_free_song",
    "bl_version", "bl_initialize_song", "bl_mean", "bl_variance", "bl_rectangular_filter",
]
BLX_H_SYMBOLS = [
    "blx_device_count", "blx_init", "blx_shutdown", "blx_last_error", "blx_configure", "blx_configure_sub_batch", "blx_debug_flags",
    "blx_analyze_batch_s16", "blx_analyze_batch_f32", "blx_analyze_batch_f32_exact", "blx_analyze_device", "blx_analyze_device_async", "blx_join",
    "blx_spectral_device",
    "blx_distance_matrix", "blx_cosine_matrix", "blx_distance_rows_device", "blx_distance_nearest_device", "blx_cosine_nearest_device",
    "blx_mean_variance_s16", "blx_rectangular_filter", "blx_frontend_f32", "blx_resample_to_s16", "blx_resample_s16_to_s16", "blx_flac_decode_frames", "blx_flac_decode_resample", "blx_flac_accelerated_count", "blx_envelope_energy_s16",
    "blx_frequency_spectrum_s16", "blx_histogram_s16", "blx_envelope_tail", "blx_envelope_energy_f32",
    "blx_profile_enable", "blx_profile_reset", "blx_profile_read", "blx_kernel_name", "blx_launch_count",
    "blx_measure_fp64_peak", "blx_multi_init", "blx_multi_shutdown", "blx_multi_device_count", "blx_multi_transport",
    "blx_multi_engine", "blx_multi_songs_taken", "blx_multi_analyze_batch_s16", "blx_multi_analyze_batch_f32", "blx_multi_set_vectors", "blx_multi_nearest",
]

_lib = None


def load():
    """Load libbliss.so and declare the prototypes. Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C bliss_b200` (or __graft_entry__.build()). "
            "bliss_b200 has no CPU implementation to fall back to.")
    L = ctypes.CDLL(LIB_PATH)
    vp = ctypes.c_void_p
    L.blx_device_count.restype = ctypes.c_int
    L.blx_init.restype = ctypes.c_int
    L.blx_init.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    L.blx_shutdown.restype = None
    L.blx_shutdown.argtypes = [vp]
    L.blx_last_error.restype = ctypes.c_char_p
    L.blx_configure.restype = ctypes.c_int
    L.blx_configure.argtypes = [vp, ctypes.c_size_t]
    L.blx_configure_sub_batch.restype = ctypes.c_int
    L.blx_configure_sub_batch.argtypes = [vp, ctypes.c_int]
    L.blx_debug_flags.restype = ctypes.c_int
    L.blx_debug_flags.argtypes = [vp, ctypes.c_uint]
    L.blx_analyze_batch_s16.restype = ctypes.c_int
    L.blx_analyze_batch_s16.argtypes = [vp, ctypes.POINTER(vp), c_i32p, c_i32p, c_u64p, ctypes.c_int,
                                        ctypes.c_uint, ctypes.POINTER(BlxResult)]
    L.blx_analyze_batch_f32.restype = ctypes.c_int
    L.blx_analyze_batch_f32.argtypes = [vp, ctypes.POINTER(vp), c_i64p, ctypes.c_int, ctypes.c_uint,
                                        ctypes.POINTER(BlxResult)]
    L.blx_analyze_batch_f32_exact.restype = ctypes.c_int
    L.blx_analyze_batch_f32_exact.argtypes = [vp, ctypes.POINTER(vp), c_i64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint,
                                              ctypes.POINTER(BlxResult)]
    L.blx_analyze_device.restype = ctypes.c_int
    L.blx_analyze_device.argtypes = [vp, ctypes.c_int, vp, c_i64p, c_i64p, c_i32p, c_u64p, ctypes.c_int,
                                     ctypes.c_uint, vp, vp]
    L.blx_analyze_device_async.restype = ctypes.c_int
    L.blx_analyze_device_async.argtypes = L.blx_analyze_device.argtypes
    L.blx_join.restype = ctypes.c_int
    L.blx_join.argtypes = [vp, vp]
    L.blx_spectral_device.restype = ctypes.c_int
    L.blx_spectral_device.argtypes = [vp, ctypes.c_int, vp, c_i64p, c_i64p, c_i32p, ctypes.c_int, vp, vp]
    for name in ("blx_distance_matrix", "blx_cosine_matrix"):
        f = getattr(L, name)
        f.restype = ctypes.c_int
        f.argtypes = [vp, c_f32p, ctypes.c_int, c_f32p]
    L.blx_distance_rows_device.restype = ctypes.c_int
    L.blx_distance_rows_device.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]
    L.blx_distance_nearest_device.restype = ctypes.c_int
    L.blx_distance_nearest_device.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
    L.blx_cosine_nearest_device.restype = ctypes.c_int
    L.blx_cosine_nearest_device.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp]
    L.blx_mean_variance_s16.restype = ctypes.c_int
    L.blx_mean_variance_s16.argtypes = [vp, c_i16p, ctypes.c_int, c_i32p, c_i32p, c_i32p]
    L.blx_rectangular_filter.restype = ctypes.c_int
    L.blx_rectangular_filter.argtypes = [vp, c_f64p, c_f64p, ctypes.c_int, ctypes.c_int]
    L.blx_frontend_f32.restype = ctypes.c_int
    L.blx_frontend_f32.argtypes = [vp, c_f32p, ctypes.c_int64, c_i16p]
    L.blx_resample_to_s16.restype = ctypes.c_int
    L.blx_resample_to_s16.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                      ctypes.c_int, c_i16p, ctypes.c_int64, c_i64p]
    L.blx_resample_s16_to_s16.restype = ctypes.c_int
    L.blx_resample_s16_to_s16.argtypes = [vp, c_i16p, ctypes.c_int, ctypes.c_int64, ctypes.c_int, c_i16p, ctypes.c_int64, c_i64p]
    L.blx_envelope_energy_s16.restype = ctypes.c_int
    L.blx_envelope_energy_s16.argtypes = [vp, c_i16p, ctypes.c_int, c_f64p]
    L.blx_envelope_energy_f32.restype = ctypes.c_int
    L.blx_envelope_energy_f32.argtypes = [vp, c_f32p, ctypes.c_int64, c_f64p]
    L.blx_frequency_spectrum_s16.restype = ctypes.c_int
    L.blx_frequency_spectrum_s16.argtypes = [vp, c_i16p, ctypes.c_int, ctypes.c_int, c_f32p]
    L.blx_histogram_s16.restype = ctypes.c_int
    L.blx_histogram_s16.argtypes = [vp, c_i16p, ctypes.c_int, ctypes.POINTER(ctypes.c_uint), c_i32p, c_i32p]
    L.blx_envelope_tail.restype = ctypes.c_int
    L.blx_envelope_tail.argtypes = [vp, c_f64p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, c_i32p, c_f32p, c_f32p]
    L.blx_profile_enable.restype = ctypes.c_int
    L.blx_profile_enable.argtypes = [vp, ctypes.c_int]
    L.blx_profile_reset.restype = ctypes.c_int
    L.blx_profile_reset.argtypes = [vp]
    L.blx_profile_read.restype = ctypes.c_int
    L.blx_profile_read.argtypes = [vp, c_f32p, c_i32p]
    L.blx_measure_fp64_peak.restype = ctypes.c_int
    L.blx_measure_fp64_peak.argtypes = [vp, c_f64p, c_f64p]
    L.blx_kernel_name.restype = ctypes.c_char_p
    L.blx_kernel_name.argtypes = [ctypes.c_int]
    L.blx_launch_count.restype = ctypes.c_longlong
    L.blx_launch_count.argtypes = [vp]
    # bliss.h
    L.bl_analyze.restype = ctypes.c_int
    L.bl_analyze.argtypes = [ctypes.c_char_p, ctypes.POINTER(BlSong)]
    L.bl_audio_decode.restype = ctypes.c_int
    L.bl_audio_decode.argtypes = [ctypes.c_char_p, ctypes.POINTER(BlSong)]
    L.bl_free_song.restype = None
    L.bl_free_song.argtypes = [ctypes.POINTER(BlSong)]
    L.bl_initialize_song.restype = None
    L.bl_initialize_song.argtypes = [ctypes.POINTER(BlSong)]
    L.bl_version.restype = ctypes.c_float
    L.bl_distance.restype = ctypes.c_float
    L.bl_distance.argtypes = [ForceVector, ForceVector]
    L.bl_cosine_similarity.restype = ctypes.c_float
    L.bl_cosine_similarity.argtypes = [ForceVector, ForceVector]
    for name in ("bl_distance_file", "bl_cosine_similarity_file"):
        f = getattr(L, name)
        f.restype = ctypes.c_float
        f.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(BlSong), ctypes.POINTER(BlSong)]
    L.bl_envelope_sort.restype = None
    L.bl_envelope_sort.argtypes = [ctypes.POINTER(BlSong), ctypes.POINTER(EnvelopeResult)]
    L.bl_amplitude_sort.restype = ctypes.c_float
    L.bl_amplitude_sort.argtypes = [ctypes.POINTER(BlSong)]
    L.bl_frequency_sort.restype = ctypes.c_float
    L.bl_frequency_sort.argtypes = [ctypes.POINTER(BlSong)]
    L.bl_mean.restype = ctypes.c_int
    L.bl_mean.argtypes = [c_i16p, ctypes.c_int]
    L.bl_variance.restype = ctypes.c_int
    L.bl_variance.argtypes = [c_i16p, ctypes.c_int, ctypes.c_int]
    L.bl_rectangular_filter.restype = None
    L.bl_rectangular_filter.argtypes = [c_f64p, c_f64p, ctypes.c_int, ctypes.c_int]
    _lib = L
    return L
