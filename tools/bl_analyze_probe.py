"""bl_analyze() from 1..16 caller threads on 3-minute WAVs and the 11-s fixture, repeated: isolates the
`bl_analyze_path` leg of bench.py (host read + decode + pageable H2D + GPU analysis per call) so its
thread scaling can be read without the rest of the bench in the process. Prints one JSON line per round."""
import ctypes
import json
import os
import struct
import sys
import tempfile
import threading
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bliss_b200  # noqa: E402

L = bliss_b200.load()
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def run(files, threads, reps):
    def work(tid):
        for _ in range(reps):
            for k in range(tid, len(files), threads):
                s = bliss_b200.BlSong()
                rc = L.bl_analyze(files[k], ctypes.byref(s))
                assert rc >= 0
                L.bl_free_song(ctypes.byref(s))
    ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return len(files) * reps / (time.perf_counter() - t0)


def main():
    rng = np.random.default_rng(7)
    n_frames = 180 * 22050
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        wavs = []
        for i in range(16):
            raw = (rng.standard_normal(2 * n_frames) * 6000).astype(np.int16).tobytes()
            path = os.path.join(d, f"s{i}.wav")
            with open(path, "wb") as f:
                f.write(b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " +
                        struct.pack("<IHHIIHH", 16, 1, 2, 22050, 22050 * 4, 4, 16) + b"data" + struct.pack("<I", len(raw)) + raw)
            wavs.append(path.encode())
        wav44 = []
        for i in range(8):  # the most common real-world input: CD audio, 44.1 kHz 16-bit stereo -> resampler -> analysis
            raw = (rng.standard_normal(2 * 180 * 44100) * 6000).astype(np.int16).tobytes()
            path = os.path.join(d, f"c{i}.wav")
            with open(path, "wb") as f:
                f.write(b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " +
                        struct.pack("<IHHIIHH", 16, 1, 2, 44100, 44100 * 4, 4, 16) + b"data" + struct.pack("<I", len(raw)) + raw)
            wav44.append(path.encode())
        fixture = os.path.join(ROOT, "tests", "golden", "song.flac").encode()
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from flac_util import repeat_flac
        long_flac = os.path.join(d, "long.flac")  # 177 s of 22 050 Hz stereo FLAC: the fixture's frames 16 times over
        with open(long_flac, "wb") as f:
            f.write(repeat_flac(fixture.decode(), 16))
        long_flac = long_flac.encode()
        # CD audio in FLAC (44.1 kHz / 16 bit / stereo, 3 minutes): 6 s from the test encoder, its frames 30 times over
        from flac_encode import encode
        from synth import song_f32
        x = song_f32(11, 6.0)
        xs = np.round(np.stack([x * 0.9, np.roll(x, 23) * 0.6], axis=1) * 32767).astype(np.int64)
        six = os.path.join(d, "six.flac")
        with open(six, "wb") as f:
            f.write(encode(xs, 16, 44100, 4096, lambda fi: dict(kind="lpc", stereo=10, lpc_order=8, porder=3), seed=3))
        cd_flac = os.path.join(d, "cd.flac")
        with open(cd_flac, "wb") as f:
            f.write(repeat_flac(six, 30))
        cd_flac = cd_flac.encode()
        run(wavs, 16, 1)
        for rnd in range(3):
            rec = {"round": rnd}
            for th in (1, 2, 4, 8, 16):
                rec[f"wav_{th}"] = round(run(wavs, th, 4), 1)
            for th in (1, 8):
                rec[f"fixture_{th}"] = round(run([fixture] * 16, th, 4), 1)
            for th in (1, 8):
                rec[f"wav44k_{th}"] = round(run(wav44, th, 2), 1)
            for th in (1, 8):
                rec[f"flac3min_gpu_{th}"] = round(run([long_flac] * 8, th, 2), 1)
            for th in (1, 8):
                rec[f"cdflac_gpu_{th}"] = round(run([cd_flac] * 8, th, 2), 1)
            os.environ["BLX_FLAC_GPU"] = "0"
            for th in (1, 8):
                rec[f"flac3min_cpu_{th}"] = round(run([long_flac] * 8, th, 2), 1)
            for th in (1, 8):
                rec[f"cdflac_cpu_{th}"] = round(run([cd_flac] * 8, th, 2), 1)
            del os.environ["BLX_FLAC_GPU"]
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
