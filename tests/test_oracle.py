"""CPU: pins the oracle (our C restatement) against the reference's golden vectors and, bit for bit,
against the reference's own analyser sources compiled verbatim (oracle/_ref)."""
import ctypes
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from synth import song_f32, song_s16

# reference tests/test_analyze.c:30-45 (song.flac) — abs tolerance 1e-5 (tests/test_analyze.c:5-11)
GOLDEN_S16 = dict(force=-20.777929, tempo=-8.945454, amplitude=-10.641844, frequency=-10.136086,
                  attack=-15.560563, nSamples=488138, duration=11, beat=59)
# reference tests/test_decode.c:16-17
GOLDEN_S16_MD5 = "8a1bd824951c0433cc47fec5bf41d0a9"
# reference tests/test_analyze.c:59-78 (audio/song_s32.flac, 48 kHz / 24 bit, through libswresample) and
# tests/test_decode.c:35-36,55-56; the fixtures are decoded by the product's FLAC reader and resampled by the
# oracle restatement of libswresample (oracle/resample.c), md5-pinned
GOLDEN_S32 = dict(force=-20.821571, tempo=-8.218182, amplitude=-10.641695, frequency=-10.179875,
                  attack=-15.561186, nSamples=488140, duration=11, beat=61)
GOLDEN_S32_MD5 = "eb9f31a7b9ed022d66ff82b76e7c3c18"
GOLDEN_S32_MONO_MD5 = "747dbfcd75bebc23ebe2024935aede36"
_S32_CACHE = {}


def load_s32_pcm(name="song_s32.flac", md5=GOLDEN_S32_MD5):
    """The reference's 48 kHz / 24-bit fixture as its decoder hands it to the analysers: int16 / 22 050 Hz / stereo."""
    if name not in _S32_CACHE:
        from flac_util import read_pcm_file
        from oracle.binding import Oracle
        orc = Oracle()
        a, n, ch, rate, bps = read_pcm_file(os.path.join(GOLDEN_DIR, name))
        assert (n, rate, bps) == (531307, 48000, 24)
        pcm = orc.resample_to_s16(a, orc.RS_S32, bps, ch, rate)
        assert pcm.dtype == np.int16 and len(pcm) == GOLDEN_S32["nSamples"]
        assert hashlib.md5(pcm.tobytes()).hexdigest() == md5  # reference tests/test_decode.c:35-36 / :55-56
        _S32_CACHE[name] = pcm
    return _S32_CACHE[name]


@pytest.fixture(scope="module")
def song_pcm():
    """PCM of the reference's audio/song.flac (copied to tests/golden), decoded by the product's host
    FLAC reader — no GPU involved."""
    import bliss_b200
    L = bliss_b200.load()
    s = bliss_b200.BlSong()
    rc = L.bl_audio_decode(os.path.join(GOLDEN_DIR, "song.flac").encode(), ctypes.byref(s))
    assert rc == 0
    pcm = np.ctypeslib.as_array(ctypes.cast(s.sample_array, ctypes.POINTER(ctypes.c_int16)), (s.nSamples,)).copy()
    meta = dict(nSamples=s.nSamples, channels=s.channels, sample_rate=s.sample_rate, duration=s.duration,
                bitrate=s.bitrate, nb_bytes_per_sample=s.nb_bytes_per_sample, artist=s.artist, title=s.title,
                album=s.album, tracknumber=s.tracknumber, genre=s.genre, resampled=s.resampled)
    L.bl_free_song(ctypes.byref(s))
    return pcm, meta


def test_decode_matches_reference_pins(song_pcm):
    pcm, meta = song_pcm
    assert hashlib.md5(pcm.tobytes()).hexdigest() == GOLDEN_S16_MD5  # reference tests/test_decode.c:16-21
    # reference tests/test_analyze.c:36-55
    assert meta["nSamples"] == 488138 and meta["channels"] == 2 and meta["sample_rate"] == 22050
    assert meta["bitrate"] == 233864 and meta["nb_bytes_per_sample"] == 2 and meta["duration"] == 11
    assert (meta["artist"], meta["title"], meta["album"], meta["tracknumber"], meta["genre"]) == (
        b"David TMX", b"Renaissance", b"Renaissance", b"02", b"Pop")


def test_oracle_reproduces_golden_vector(oracle, song_pcm):
    pcm, _ = song_pcm
    r = oracle.analyze(pcm, GOLDEN_S16["duration"])
    for k in ("force", "tempo", "amplitude", "frequency", "attack"):
        assert abs(r[k] - GOLDEN_S16[k]) <= 1e-5, (k, r[k], GOLDEN_S16[k])
    assert r["beat"] == GOLDEN_S16["beat"]
    assert r["calm_or_loud"] == 1


def test_reference_sources_reproduce_golden_vector(reflib, song_pcm):
    pcm, _ = song_pcm
    r = reflib.bl_analyze(pcm, GOLDEN_S16["duration"])
    assert r["rc"] == 1
    for k in ("force", "tempo", "amplitude", "frequency", "attack"):
        assert abs(r[k] - GOLDEN_S16[k]) <= 1e-5, (k, r[k], GOLDEN_S16[k])


def test_second_golden_vector_s32(oracle, reflib):
    """The reference's second fixture: both checkers reproduce its pinned force vector (abs 1e-5, reference
    tests/test_analyze.c:5-11) from the PCM the reference's decoder produces (md5-pinned)."""
    pcm = load_s32_pcm()
    for r in (oracle.analyze(pcm, GOLDEN_S32["duration"]), reflib.bl_analyze(pcm, GOLDEN_S32["duration"])):
        for k in ("force", "tempo", "amplitude", "frequency", "attack"):
            assert abs(r[k] - GOLDEN_S32[k]) <= 1e-5, (k, r[k], GOLDEN_S32[k])
    assert oracle.analyze(pcm, GOLDEN_S32["duration"])["beat"] == GOLDEN_S32["beat"]


def _bits(x):
    return np.float32(x).tobytes()


@pytest.mark.parametrize("seed", range(6))
def test_oracle_bit_identical_to_reference_sources(oracle, reflib, seed):
    seconds = [3.0, 5.5, 8.0, 2.2, 12.0, 4.1][seed]
    pcm = song_s16(seed, seconds, decorrelate=bool(seed % 2), gain=[1.0, 0.3, 0.05, 1.5, 0.8, 0.01][seed])
    dur = max(1, int(seconds))
    a = oracle.analyze(pcm, dur)
    b = reflib.bl_analyze(pcm, dur)
    for k in ("tempo", "amplitude", "frequency", "attack", "force"):
        assert _bits(a[k]) == _bits(b[k]), (k, a[k], b[k])
    assert a["calm_or_loud"] == b["calm_or_loud"]


def test_oracle_silence_trimming_matches_reference(oracle, reflib):
    """Zeros in front of the first and behind the last non-zero sample are skipped by the amplitude
    analyser (reference src/amplitude_sort.c:26-39); silence inside the song is not."""
    pcm = song_s16(13, 6.0, decorrelate=True)
    pcm[:5001] = 0
    pcm[-3000:] = 0
    pcm[100000:160000] = 0
    a, b = oracle.analyze(pcm, 6), reflib.bl_analyze(pcm, 6)
    for k in ("tempo", "amplitude", "frequency", "attack", "force"):
        assert _bits(a[k]) == _bits(b[k]), (k, a[k], b[k])


def test_oracle_mono_branch_matches_reference(oracle, reflib):
    pcm = song_s16(11, 3.0)[::2].copy()  # reference src/frequency_sort.c:76-80
    s, keep = reflib.make_song(pcm, 3, channels=1)
    assert _bits(oracle.frequency(pcm, channels=1)) == _bits(reflib.lib.bl_frequency_sort(ctypes.byref(s)))


def test_oracle_helpers_match_reference(oracle, reflib):
    pcm = song_s16(3, 2.0, decorrelate=True)
    m, v = oracle.mean_variance(pcm)
    p = pcm.ctypes.data_as(ctypes.POINTER(ctypes.c_int16))
    assert m == reflib.lib.bl_mean(p, len(pcm))
    assert v == reflib.lib.bl_variance(p, len(pcm), m)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(500)
    out0 = rng.standard_normal(500)
    ours = oracle.rectangular_filter(out0, x, 19)
    ref = out0.copy()
    reflib.lib.bl_rectangular_filter(ref.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                     x.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 500, 19)
    assert np.array_equal(ours, ref)


def test_oracle_distance_matches_reference(oracle, reflib):
    rng = np.random.default_rng(7)
    v = (rng.standard_normal((40, 4)) * 10).astype(np.float32)
    for i in range(0, 40, 3):
        for j in range(0, 40, 5):
            assert _bits(oracle.distance(v[i], v[j])) == _bits(reflib.distance(v[i], v[j]))
            if i != j:
                assert _bits(oracle.cosine_similarity(v[i], v[j])) == _bits(reflib.cosine_similarity(v[i], v[j]))


def test_frontend_spec_properties(oracle):
    x = song_f32(1, 1.0)
    q = oracle.frontend_f32(x)
    assert len(q) == 2 * (len(x) // 2) and np.array_equal(q[0::2], q[1::2])
    # DC gain = 32768 * 0.70710678 (blx_frontend.h)
    dc = oracle.frontend_f32(np.full(4000, 0.25, dtype=np.float32))
    assert abs(int(dc[2000]) - round(0.25 * 32768 * 0.70710678)) <= 1
    # clipping
    big = oracle.frontend_f32(np.full(4000, 4.0, dtype=np.float32))
    assert big[2000] == 32767


def test_chain_model_against_sequential_chain(tmp_path):
    """tools/chain_model.c: the predicted-binade accumulation of csrc/envelope.cu (float_chain) restated on
    the host, against the sequential chain s <- (float)((double)s + p_k) of reference
    src/tempo_atk_sort.c:142-150 on random spectra (flat, coloured, tonal, stepped, wide dynamic range)."""
    import subprocess
    exe = tmp_path / "chain_model"
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "chain_model.c")
    subprocess.run(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-o", str(exe), src, "-lm"], check=True)
    out = subprocess.run([str(exe), "120000"], check=True, capture_output=True, text=True).stdout
    last = out.strip().splitlines()[-1].split()
    stats = dict(zip(last[0::2], last[1::2]))
    assert stats["mismatches"] == "0", out
    assert int(stats["fallback"]) < 120000 * 1e-3, out  # verified fast path on > 99.9 % of the spectra


# ---------------------------------------------------------------- decode-stage resampler (include/blx_resample.h)
def test_resampler_reproduces_reference_md5_pins():
    """reference tests/test_decode.c:27-65: both 48 kHz fixtures through our FLAC reader and the oracle restatement of
    libswresample give the reference's md5 (load_s32_pcm asserts it), the mono one up-mixed to L == R."""
    load_s32_pcm()
    mono = load_s32_pcm("song_s32_mono.flac", GOLDEN_S32_MONO_MD5)
    assert np.array_equal(mono[0::2], mono[1::2])


def test_resampler_matches_libswresample_vectors(oracle):
    """tests/golden/resample_vectors.npz (tools/make_golden_resample.py): seeded inputs and what libswresample 6.1.100
    made of them - every rate / sample format / channel count bit for bit."""
    z = np.load(os.path.join(GOLDEN_DIR, "resample_vectors.npz"))
    keys = sorted(k[:-3] for k in z.files if k.endswith("_in"))
    assert len(keys) >= 20
    kinds = dict(s16=oracle.RS_S16, s32=oracle.RS_S32, f32=oracle.RS_F32, u8=oracle.RS_U8)
    for key in keys:
        kind, bits, ch, rate = key.split("_")
        got = oracle.resample_to_s16(z[key + "_in"], kinds[kind], int(bits), int(ch), int(rate))
        assert np.array_equal(got, z[key + "_out"]), key


def test_resampler_plan_properties(oracle):
    """DC gain 1 in every phase, linear phase, and the frame counts libswresample produces."""
    for rate, frames, want in ((48000, 531307, 244070), (44100, 20001, 10001), (8000, 9999, 27560), (96000, 13844, 3180)):
        x = np.zeros(frames * 2, dtype=np.int32)
        assert len(oracle.resample_to_s16(x, oracle.RS_S16, 16, 2, rate)) == 2 * want, rate
    dc = np.full(2 * 6000, 12345, dtype=np.int32)
    for rate in (48000, 44100, 32000, 11025):
        y = oracle.resample_to_s16(dc, oracle.RS_S16, 16, 2, rate)
        assert np.all(np.abs(y.astype(np.int32) - 12345) <= 1), rate  # mirrored ends: no edge transient either
