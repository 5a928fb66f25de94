"""bliss_b200 — B200-native drop-in for the bliss music analyser's hot path.

The package is a thin ctypes face over the in-tree libbliss.so (host C + hand-written sm_100a
CUDA). Importing it never falls back to a CPU implementation: without the built library
`load()` raises, without a B200 `Engine()` raises.
"""
from ._lib import LIB_PATH, BlSong, BlxResult, EnvelopeResult, ForceVector, load  # noqa: F401
from . import compat  # noqa: F401  (the reference's Python package surface)
from .engine import (ALIGN_ELEMS, DO_ALL, DO_AMPLITUDE, DO_ENVELOPE, DO_FREQUENCY, FMT_F32, FMT_S16,  # noqa: F401
                     RESULT_DTYPE, BlxError, Engine)

__version__ = "0.1.0"
