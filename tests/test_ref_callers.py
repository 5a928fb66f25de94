"""GPU: the drop-in boundary proven with the REFERENCE'S OWN callers (not ctypes):
  - reference tests/test_analyze.c, examples/analyze.c, examples/distance.c compiled VERBATIM from /root/reference
    against include/bliss.h + bliss_b200/libbliss.so (oracle/Makefile `callers`, built in the authoring container,
    shipped prebuilt under oracle/_ref/bin);
  - the reference's cffi module built per reference python/build_bliss.py:21-38 but re-pointed at libbliss.so
    (tools/build_ref_cffi.py) with the reference's UNMODIFIED Python package (bl_song.py, distance.py, version.py).
"""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT
from test_oracle import GOLDEN_S16, GOLDEN_S32

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
PYREF = os.path.join(ROOT, "oracle", "_ref", "pyref")


def _need(path):
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (needs /root/reference at build time: __graft_entry__.build())")
    return path


@pytest.fixture(scope="module")
def audio_dir(tmp_path_factory):
    """<tmp>/audio with the reference's fixtures and <tmp>/run as working directory: the test programme opens "../audio/..."."""
    base = tmp_path_factory.mktemp("refcallers")
    os.makedirs(base / "audio")
    os.makedirs(base / "run")
    for f in ("song.flac", "song_s32.flac", "song_s32_mono.flac"):
        shutil.copyfile(os.path.join(GOLDEN_DIR, f), base / "audio" / f)
    return base


def test_reference_test_analyze_c_passes_verbatim(audio_dir):
    """reference tests/test_analyze.c: test_analyze_s16 (song.flac, native) AND test_analyze_s32 (48 kHz / 24 bit through the
    resampler): force vector to 1e-5, nSamples, bitrate, duration, tags. Any failed assertion exits with -1."""
    exe = _need(os.path.join(BIN, "test_analyze"))
    r = subprocess.run([exe], cwd=audio_dir / "run", capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)


def test_reference_example_analyze(audio_dir):
    exe = _need(os.path.join(BIN, "example_analyze"))
    r = subprocess.run([exe, str(audio_dir / "audio" / "song.flac")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert "Force: %f" % GOLDEN_S16["force"] in out or "Force: -20.7779" in out
    assert "Number of samples: 488138" in out and "Calm or loud: Calm" in out and "Duration: 11" in out
    assert "Artist: David TMX" in out and "Track number: 02" in out and "Genre: Pop" in out
    vec = out.split("Force vector: (")[1].split(")")[0].split(", ")
    for v, k in zip(vec, ("tempo", "amplitude", "frequency", "attack")):
        assert abs(float(v) - GOLDEN_S16[k]) <= 2e-6 + 1e-5, (k, v)
    bad = subprocess.run([exe, str(audio_dir / "audio" / "missing.flac")], capture_output=True, text=True, timeout=300)
    assert bad.returncode != 0 and "Couldn't analyze song" in bad.stderr


def test_reference_example_distance(audio_dir, oracle):
    exe = _need(os.path.join(BIN, "example_distance"))
    f1, f2 = str(audio_dir / "audio" / "song.flac"), str(audio_dir / "audio" / "song_s32.flac")
    r = subprocess.run([exe, f1, f2], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    a = [GOLDEN_S16[k] for k in ("tempo", "amplitude", "frequency", "attack")]
    b = [GOLDEN_S32[k] for k in ("tempo", "amplitude", "frequency", "attack")]
    d = float(r.stdout.split("is: ")[1].split()[0])
    c = float(r.stdout.split("is: ")[2].split()[0])
    assert abs(d - oracle.distance(a, b)) <= 2e-5 and d > 0.5  # two different songs: not the trivial 0
    assert abs(c - oracle.cosine_similarity(a, b)) <= 1e-5


def test_reference_python_package_on_libbliss(audio_dir):
    """The reference's python/bliss package, unmodified, over the re-pointed cffi module."""
    _need(os.path.join(PYREF, "bliss", "bl_song.py"))
    code = r"""
import sys
import bliss
from bliss import bl_song, distance, version
f1, f2 = sys.argv[1], sys.argv[2]
with bl_song(f1) as song:
    fv = song["force_vector"]
    print("FV", fv["tempo"], fv["amplitude"], fv["frequency"], fv["attack"], song["force"], song["calm_or_loud"], song["nSamples"],
          song["duration"], song["artist"], song["title"], song["tracknumber"])
    env = song.envelope_analysis()
    # (the binding's amplitude_analysis / frequency_analysis call the library and return None: reference bl_song.py:184-199)
    print("ENV", env["tempo"], env["attack"], song.amplitude_analysis(), song.frequency_analysis())
d = distance.distance(f1, f2)
print("DIST", d["distance"], d["song2"]["nSamples"], d["song2"]["force"])
d["song1"].free(); d["song2"].free()
print("VERSION", version.version())
"""
    env = dict(os.environ, PYTHONPATH=PYREF)
    r = subprocess.run([sys.executable, "-c", code, str(audio_dir / "audio" / "song.flac"), str(audio_dir / "audio" / "song_s32.flac")],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    lines = {ln.split()[0]: ln.split()[1:] for ln in r.stdout.splitlines() if ln and ln.split()[0] in ("FV", "ENV", "DIST", "VERSION")}
    fv = [float(x) for x in lines["FV"][:5]]
    for v, k in zip(fv, ("tempo", "amplitude", "frequency", "attack", "force")):
        assert abs(v - GOLDEN_S16[k]) <= 1e-5, (k, v)
    assert lines["FV"][5:8] == ["1", "488138", "11"] and "David" in lines["FV"][8]
    assert [float(x) for x in lines["ENV"][:2]] == [fv[0], fv[3]] and lines["ENV"][2:] == ["None", "None"]
    assert float(lines["DIST"][0]) > 0.5 and lines["DIST"][1] == "488140" and abs(float(lines["DIST"][2]) - GOLDEN_S32["force"]) <= 1e-5
    assert abs(float(lines["VERSION"][0]) - 1.2) < 1e-6
