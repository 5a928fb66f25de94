"""The host file reader parses untrusted files (bliss_b200/host/flac_reader.c): a short mutation-fuzzing run under
AddressSanitizer + UBSan (tools/fuzz_reader.c; 30 000 mutated files were run once by hand). Skipped
where gcc has no sanitizer runtime."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reader_survives_mutated_files(tmp_path):
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    exe = tmp_path / "fuzz_reader"
    cmd = [gcc, "-std=gnu99", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
           "-I" + os.path.join(ROOT, "bliss_b200", "host"), "-o", str(exe), os.path.join(ROOT, "tools", "fuzz_reader.c"),
           os.path.join(ROOT, "bliss_b200", "host", "flac_reader.c"), "-lm"]
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("gcc cannot link the sanitizer runtimes here: " + build.stderr[-200:])
    rng = np.random.default_rng(3)
    raw = (rng.standard_normal(22050 * 2) * 3000).astype(np.int16).tobytes()
    wav16 = tmp_path / "a.wav"
    wav16.write_bytes(b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 2, 22050, 88200, 4, 16) +
                      b"data" + struct.pack("<I", len(raw)) + raw)
    raw = (rng.standard_normal(24000) * 0.2).astype(np.float32).tobytes()
    wavf = tmp_path / "b.wav"
    wavf.write_bytes(b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 3, 1, 48000, 192000, 4, 32) +
                     b"data" + struct.pack("<I", len(raw)) + raw)
    seeds = [os.path.join(ROOT, "tests", "golden", n) for n in ("song.flac", "song_s32_mono.flac")] + [str(wav16), str(wavf)]
    run = subprocess.run([str(exe), "400"] + seeds, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    assert "no memory or undefined-behaviour error" in run.stdout
