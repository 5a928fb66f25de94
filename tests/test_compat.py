"""The reference's Python package surface (reference python/bliss/*.py) over the B200 library:
bliss_b200.compat. CPU tests use the host-only entry points (decode, by-value distances); the GPU tests
analyse the reference's fixture and build a playlist."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from test_oracle import GOLDEN_S16

FLAC = os.path.join(GOLDEN_DIR, "song.flac")


def test_bl_song_mapping_and_decode():
    from bliss_b200 import compat as bliss
    assert (bliss.BL_LOUD, bliss.BL_CALM, bliss.BL_UNKNOWN, bliss.BL_UNEXPECTED, bliss.BL_OK) == (0, 1, 2, -2, 0)
    with bliss.bl_song() as song:
        assert len(song) == 17 and "force_vector" in list(song)
        assert song["title"] is None and song["sample_array"] == []
        song["title"] = "a title"
        song["force_vector"] = {"tempo": 1.0, "amplitude": 2.0, "frequency": 3.0, "attack": 4.0}
        song["force"] = 10.0
        assert song["title"] == "a title" and song["force_vector"]["frequency"] == 3.0 and song["force"] == 10.0
        with pytest.raises(KeyError):
            song["nope"]
    with bliss.bl_song() as song:  # reference tests/test_decode.c:12-27: nSamples, rate, channels of the fixture
        assert song.decode(FLAC) == bliss.BL_OK
        assert song["nSamples"] == 488138 and song["sample_rate"] == 22050 and song["channels"] == 2
        assert song["duration"] == 11 and len(song["sample_array"]) == 488138
        assert isinstance(song["title"], str) and isinstance(song["artist"], str)
    assert isinstance(bliss.version(), float)


def test_distance_wrappers_by_value(oracle):
    from bliss_b200 import compat as bliss
    a, b = bliss.bl_song(), bliss.bl_song()
    a["force_vector"] = (-8.9, -10.6, -10.1, -15.5)
    b["force_vector"] = {"tempo": 2.5, "amplitude": 1.25, "frequency": -3.0, "attack": 7.75}
    va = np.array([-8.9, -10.6, -10.1, -15.5], dtype=np.float32)
    vb = np.array([2.5, 1.25, -3.0, 7.75], dtype=np.float32)
    d = bliss.distance(a, b)
    assert d["song1"] is a and d["song2"] is b
    assert np.float32(d["distance"]) == np.float32(oracle.distance(va, vb))
    c = bliss.cosine_similarity(a, b)
    assert np.float32(c["similarity"]) == np.float32(oracle.cosine_similarity(va, vb))
    assert bliss.distance(a, "x.flac") == {"distance": None, "song1": None, "song2": None}
    a.free()
    b.free()


@pytest.mark.gpu
def test_bl_song_analyze_fixture_and_file_distance():
    from bliss_b200 import compat as bliss
    with bliss.bl_song(FLAC) as song:  # reference tests/test_analyze.c:26-57
        fv = song["force_vector"]
        for k in ("tempo", "amplitude", "frequency", "attack"):
            assert abs(fv[k] - GOLDEN_S16[k]) <= 1e-5, (k, fv[k])
        assert abs(song["force"] - GOLDEN_S16["force"]) <= 1e-5 and song["calm_or_loud"] == bliss.BL_CALM
        env = song.envelope_analysis()
        assert env["tempo"] == fv["tempo"] and env["attack"] == fv["attack"]
        assert song.amplitude_analysis() == fv["amplitude"] and song.frequency_analysis() == fv["frequency"]
    d = bliss.distance(FLAC, FLAC)
    assert d["distance"] == 0.0 and d["song1"]["nSamples"] == 488138
    c = bliss.cosine_similarity(FLAC, FLAC)
    assert abs(c["similarity"] - 1.0) <= 1e-6
    for r in (d, c):
        r["song1"].free()
        r["song2"].free()
    missing = bliss.distance("/nonexistent/a.flac", FLAC)
    assert missing["distance"] == float(bliss.BL_UNEXPECTED)


@pytest.mark.gpu
def test_playlist_matches_reference_example(engine, oracle):
    """reference python/examples/make_m3u_playlist.py:62-67: euclidean distance seed -> all, argsort."""
    from bliss_b200 import compat as bliss
    rng = np.random.default_rng(5)
    v = (rng.standard_normal((5000, 4)) * np.array([6, 4, 9, 7])).astype(np.float32)
    v[77] = v[3000]
    idx, dist = bliss.playlist(engine, v, 3000)
    full = np.array([oracle.distance(v[3000], v[i]) for i in range(5000)], dtype=np.float32)
    assert np.array_equal(dist, np.sort(full, kind="stable"))
    assert np.array_equal(idx, np.argsort(full, kind="stable"))
    assert idx[0] == 77 and idx[1] == 3000 and dist[1] == 0.0  # the collision and the seed, index order
    top, _ = bliss.playlist(engine, v, 3000, k=10)
    assert np.array_equal(top, idx[:10])
