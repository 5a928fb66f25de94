"""Decode-only timing of a long FLAC stream: host threads against the device decoder (csrc/flacdec.cu)."""
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from flac_util import read_pcm_file, repeat_flac  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 16
path = "/dev/shm/blx_long.flac" if os.path.isdir("/dev/shm") else "/tmp/blx_long.flac"
open(path, "wb").write(repeat_flac(os.path.join(ROOT, "tests", "golden", "song.flac"), reps))
for mode, env in (("host x1", {"BLX_FLAC_GPU": "0", "BLX_DECODE_THREADS": "1"}), ("host x4", {"BLX_FLAC_GPU": "0", "BLX_DECODE_THREADS": "4"}),
                  ("host x8", {"BLX_FLAC_GPU": "0", "BLX_DECODE_THREADS": "8"}), ("device", {"BLX_FLAC_GPU": "1", "BLX_FLAC_GPU_MIN_SAMPLES": "0"})):
    for k in ("BLX_FLAC_GPU", "BLX_DECODE_THREADS", "BLX_FLAC_GPU_MIN_SAMPLES"):
        os.environ.pop(k, None)
    os.environ.update(env)
    read_pcm_file(path)
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        a = read_pcm_file(path)
    dt = (time.perf_counter() - t0) / n
    print(f"{mode}: {dt * 1e3:.1f} ms per decode of {a[1] * a[2] / 1e6:.1f} M samples (incl. file read and the int32 copy of the test helper)")
os.remove(path)
