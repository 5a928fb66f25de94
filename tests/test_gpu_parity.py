"""GPU: parity of the CUDA path against the oracle, through the C-ABI (include/blx.h) and the
bliss.h wrappers. Tolerances (BASELINE.json north_star): every force_vector component within 1e-4
relative of the reference, onset counts exact, distances bit-exact."""
import ctypes
import os

import numpy as np
import pytest

import bliss_b200
from conftest import GOLDEN_DIR
from synth import song_f32, song_s16
from test_oracle import GOLDEN_S16

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4  # north_star: "every force_vector within 1e-4 relative of the reference"


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-30)


def check_song(res, ref, tag=""):
    assert res["status"] == 0, (tag, res)
    assert int(res["beat"]) == ref["beat"], (tag, "beat", int(res["beat"]), ref["beat"])
    for k in ("tempo", "amplitude", "frequency", "attack", "force"):
        if np.isnan(ref[k]) and np.isnan(float(res[k])):
            continue  # e.g. anti-phase stereo: the mono mix is all zeros, 0 / 0 in the reference's dB step too
        assert rel(float(res[k]), ref[k]) <= REL_TOL, (tag, k, float(res[k]), ref[k])
    assert int(res["calm_or_loud"]) == ref["calm_or_loud"], tag


# ---------------------------------------------------------------- config 1: the reference's own fixture
def test_bl_analyze_golden_fixture(oracle):
    L = bliss_b200.load()
    s = bliss_b200.BlSong()
    rc = L.bl_analyze(os.path.join(GOLDEN_DIR, "song.flac").encode(), ctypes.byref(s))
    assert rc == 1  # BL_CALM
    got = dict(force=s.force, tempo=s.force_vector.tempo, amplitude=s.force_vector.amplitude,
               frequency=s.force_vector.frequency, attack=s.force_vector.attack)
    # reference tests/test_analyze.c:30-35 with its own 1e-5 absolute tolerance (:5-11)
    for k, v in got.items():
        assert abs(v - GOLDEN_S16[k]) <= 1e-5, (k, v, GOLDEN_S16[k])
    assert s.nSamples == 488138 and s.channels == 2 and s.duration == 11 and s.calm_or_loud == 1
    # tempo = 4 * beat / duration - 30.4 with beat = 59 exactly
    assert got["tempo"] == np.float32(np.float32(4 * np.float32(59) / np.float32(11)) - 30.4)
    pcm = np.ctypeslib.as_array(ctypes.cast(s.sample_array, ctypes.POINTER(ctypes.c_int16)), (s.nSamples,)).copy()
    # the three stand-alone analysers on the decoded song (reference include/bliss.h:184-217)
    env = bliss_b200.EnvelopeResult()
    L.bl_envelope_sort(ctypes.byref(s), ctypes.byref(env))
    assert env.tempo == got["tempo"] and env.attack == got["attack"]
    assert L.bl_amplitude_sort(ctypes.byref(s)) == got["amplitude"]
    assert L.bl_frequency_sort(ctypes.byref(s)) == got["frequency"]
    L.bl_free_song(ctypes.byref(s))
    ref = oracle.analyze(pcm, 11)
    assert got["amplitude"] == ref["amplitude"]  # histogram smoothing is replayed bit for bit
    assert rel(got["frequency"], ref["frequency"]) <= 1e-5


def test_second_golden_fixture_s32(oracle):
    """reference tests/test_analyze.c:59-90 verbatim: bl_analyze on audio/song_s32.flac (48 kHz / 24 bit) - decoded by the
    host FLAC reader, resampled on the GPU (include/blx_resample.h) - gives the reference's force vector to its own 1e-5,
    its field values, and the md5 of the resampled PCM that reference tests/test_decode.c:35-36 pins; same for the mono file."""
    import hashlib
    from test_oracle import GOLDEN_S32, GOLDEN_S32_MD5, GOLDEN_S32_MONO_MD5
    L = bliss_b200.load()
    for name, md5 in (("song_s32.flac", GOLDEN_S32_MD5), ("song_s32_mono.flac", GOLDEN_S32_MONO_MD5)):
        s = bliss_b200.BlSong()
        rc = L.bl_analyze(os.path.join(GOLDEN_DIR, name).encode(), ctypes.byref(s))
        assert rc == 1, name  # BL_CALM
        pcm = np.ctypeslib.as_array(ctypes.cast(s.sample_array, ctypes.POINTER(ctypes.c_int16)), (s.nSamples,)).copy()
        assert hashlib.md5(pcm.tobytes()).hexdigest() == md5, name
        assert (s.nSamples, s.channels, s.sample_rate, s.nb_bytes_per_sample, s.duration, s.resampled) == (488140, 2, 22050, 2, 11, 1)
        assert (s.artist, s.title, s.album, s.tracknumber, s.genre) == (b"David TMX", b"Renaissance", b"Renaissance", b"02", b"Pop")
        if name == "song_s32.flac":
            assert s.bitrate == 840742  # reference tests/test_analyze.c:74
            got = dict(force=s.force, tempo=s.force_vector.tempo, amplitude=s.force_vector.amplitude,
                       frequency=s.force_vector.frequency, attack=s.force_vector.attack)
            for k, v in got.items():
                assert abs(v - GOLDEN_S32[k]) <= 1e-5, (k, v, GOLDEN_S32[k])
        else:
            ref = oracle.analyze(pcm, 11)
            assert int(round((s.force_vector.tempo + 30.4) * 11 / 4)) == ref["beat"] and rel(s.force, ref["force"]) <= REL_TOL
        L.bl_free_song(ctypes.byref(s))


def test_resampler_kernel_matches_oracle(engine, oracle):
    """blx_resample_to_s16 against the oracle on the libswresample fixtures and on seeded signals: bit-exact."""
    z = np.load(os.path.join(GOLDEN_DIR, "resample_vectors.npz"))
    kinds = dict(s16=engine.RS_S16, s32=engine.RS_S32, f32=engine.RS_F32, u8=engine.RS_U8)
    for key in sorted(k[:-3] for k in z.files if k.endswith("_in")):
        kind, bits, ch, rate = key.split("_")
        got = engine.resample_to_s16(z[key + "_in"], kinds[kind], int(bits), int(ch), int(rate))
        assert np.array_equal(got, z[key + "_out"]), key
    rng = np.random.default_rng(77)
    for rate, ch, n in ((48000, 2, 100003), (32000, 1, 50000), (12000, 2, 30000), (192000, 2, 90001), (7350, 1, 9000)):
        x = (rng.standard_normal(n * ch) * 7000).clip(-32768, 32767).astype(np.int32)
        ref = oracle.resample_to_s16(x, oracle.RS_S16, 16, ch, rate)
        assert np.array_equal(engine.resample_to_s16(x, engine.RS_S16, 16, ch, rate), ref), rate
        assert np.array_equal(engine.resample_s16_to_s16(x.astype(np.int16), ch, rate), ref), rate  # int16 storage, same result
    for ch in (1, 2):  # same rate: up-mix / pass-through only
        x = (rng.standard_normal(5000 * ch) * 7000).clip(-32768, 32767).astype(np.int32)
        assert np.array_equal(engine.resample_s16_to_s16(x.astype(np.int16), ch, 22050), oracle.resample_to_s16(x, oracle.RS_S16, 16, ch, 22050))
    with pytest.raises(bliss_b200.BlxError):
        engine.resample_to_s16(np.zeros(4000, np.int32), engine.RS_S16, 16, 2, 47999)  # 22050 / 47999: more than 1024 phases


# ---------------------------------------------------------------- native int16 input
S16_CASES = [(0, 6.0, False, 1.0), (1, 11.3, True, 0.4), (2, 3.0, True, 0.02), (3, 20.0, False, 1.6),
             (4, 2.0, True, 0.9), (5, 33.0, False, 0.7)]


def test_batch_s16_matches_oracle(engine, oracle):
    songs = [song_s16(seed, sec, decorrelate=dec, gain=g) for seed, sec, dec, g in S16_CASES]
    durs = [max(1, int(sec)) for _, sec, _, _ in S16_CASES]
    res = engine.analyze_s16(songs, durs)
    for i, pcm in enumerate(songs):
        ref = oracle.analyze(pcm, durs[i])
        check_song(res[i], ref, tag=f"s16 case {i}")
        assert float(res[i]["amplitude"]) == ref["amplitude"], i  # bit-exact stage


def test_ragged_batch_equals_single_song_calls(engine):
    songs = [song_s16(20 + i, sec, decorrelate=bool(i & 1)) for i, sec in enumerate([2.0, 7.7, 3.3, 15.0, 2.5])]
    durs = [2, 7, 3, 15, 2]
    batch = engine.analyze_s16(songs, durs)
    for i in range(len(songs)):
        one = engine.analyze_s16([songs[i]], [durs[i]])
        assert batch[i].tobytes() == one[0].tobytes(), i  # deterministic, independent of batching


def test_long_song_is_independent_of_batch_size(engine):
    # the cut of a song into pass-1 parts (and so the float summation order of its spectrum) must
    # depend on the song alone: alone, or as one of 48 songs, the record is byte-identical
    long_song = song_s16(77, 45.0, decorrelate=True)
    alone = engine.analyze_s16([long_song], [45])
    fill = [song_s16(300 + i, 2.0) for i in range(47)]
    batch = engine.analyze_s16(fill[:20] + [long_song] + fill[20:], [2] * 20 + [45] + [2] * 27)
    assert batch[20].tobytes() == alone[0].tobytes()


def test_small_chunks_equal_one_chunk(engine):
    songs = [song_s16(40 + i, 2.0 + 0.7 * i) for i in range(6)]
    durs = [2] * 6
    a = engine.analyze_s16(songs, durs)
    small = bliss_b200.Engine(0, chunk_bytes=1 << 20)  # forces one song per chunk, both slots in rotation
    b = small.analyze_s16(songs, durs)
    small.close()
    assert a.tobytes() == b.tobytes()


def test_mono_branch(engine, oracle):
    pcm = song_s16(9, 5.0)[::2].copy()
    res = engine.analyze_s16([pcm], [5], channels=[1], what=bliss_b200.DO_FREQUENCY)
    assert rel(float(res[0]["frequency"]), oracle.frequency(pcm, channels=1)) <= 1e-5


def test_envelope_intermediates(engine, oracle):
    pcm = song_s16(7, 9.0, decorrelate=True, gain=0.5)
    m, v = engine.mean_variance(pcm)
    assert (m, v) == oracle.mean_variance(pcm)
    assert engine.mean_variance(pcm, mean_in=m + 3)[1] == oracle.lib.orc_variance(
        pcm.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)), len(pcm), m + 3)
    E = engine.envelope_energy(pcm)
    Eo = oracle.envelope_energy(pcm)
    assert E.shape == Eo.shape and np.all(E[-2:] == 0)
    # E[m] is a float-rounded sum: equal to the oracle's up to rare single-ulp (6e-8) flips
    assert np.max(np.abs(E - Eo) / np.maximum(Eo, 1e-300)) <= 2.4e-7
    assert np.mean(E == Eo) > 0.99


def test_rectangular_filter_helper(engine, oracle):
    rng = np.random.default_rng(3)
    x, out0 = rng.standard_normal(700), rng.standard_normal(700)
    assert np.array_equal(engine.rectangular_filter(out0, x, 19), oracle.rectangular_filter(out0, x, 19))


# ---------------------------------------------------------------- 44.1 kHz float32 input (front-end)
def test_frontend_int16_stream_is_bit_exact(engine, oracle):
    for seed, sec in [(0, 1.0), (1, 2.37), (2, 0.5)]:
        x = song_f32(seed, sec)
        if seed == 1:
            x = (x * 3.9).astype(np.float32)  # drive into clipping
        if seed == 2:
            x = x[:-1]  # odd length
        assert np.array_equal(engine.frontend_f32(x), oracle.frontend_f32(x)), seed


def test_batch_f32_matches_oracle(engine, oracle):
    secs = [4.0, 9.5, 30.0, 2.0]
    songs = [song_f32(100 + i, s) for i, s in enumerate(secs)]
    res = engine.analyze_f32(songs)
    for i, x in enumerate(songs):
        pcm = oracle.frontend_f32(x)  # identical int16 on both sides (blx_frontend.h)
        ref = oracle.analyze(pcm, len(x) // 44100)
        check_song(res[i], ref, tag=f"f32 case {i}")
        assert float(res[i]["amplitude"]) == ref["amplitude"], i
        # the fused path and the native-int16 path see the same samples
        again = engine.analyze_s16([pcm], [len(x) // 44100])
        assert again[0].tobytes() == res[i].tobytes(), i


# ---------------------------------------------------------------- degenerate inputs
def test_degenerate_inputs_are_flagged(engine):
    silent = np.zeros(60000, dtype=np.int16)
    short = song_s16(1, 0.05)
    flat = np.full(60000, 7, dtype=np.int16)
    ok = song_s16(2, 2.0)
    res = engine.analyze_s16([silent, short, flat, ok], [1, 1, 1, 2])
    assert res[0]["status"] & bliss_b200.engine.SONG_SILENT and np.isnan(res[0]["amplitude"])
    assert res[1]["status"] & bliss_b200.engine.SONG_TOO_SHORT and np.isnan(res[1]["tempo"])
    assert res[2]["status"] & bliss_b200.engine.SONG_FLAT
    assert res[3]["status"] == 0
    with pytest.raises(bliss_b200.BlxError):
        engine.analyze_s16([ok], [2], channels=[3])


# ---------------------------------------------------------------- distances
def test_distance_matrix_bit_exact(engine, oracle):
    rng = np.random.default_rng(11)
    v = (rng.standard_normal((301, 4)) * np.array([10, 8, 12, 15])).astype(np.float32)
    v[5] = v[17]  # a collision: distance exactly 0
    d = engine.distance_matrix(v)
    assert np.array_equal(d, oracle.distance_matrix(v))
    assert d[5, 17] == 0 and np.array_equal(d, d.T) and np.all(np.diag(d) == 0)
    c = engine.distance_matrix(v, cosine=True)
    for i in range(0, 301, 13):
        for j in range(0, 301, 7):
            assert c[i, j] == np.float32(oracle.cosine_similarity(v[i], v[j])), (i, j)


def test_distance_nearest_matches_matrix(engine):
    import torch
    rng = np.random.default_rng(12)
    v = (rng.standard_normal((1000, 4)) * 10).astype(np.float32)
    d = engine.distance_matrix(v).astype(np.float64)
    dv = torch.from_numpy(v).cuda()
    idx = torch.empty(1000, dtype=torch.int32, device="cuda")
    dist = torch.empty(1000, dtype=torch.float32, device="cuda")
    rsum = torch.empty(1000, dtype=torch.float64, device="cuda")
    engine.distance_nearest_device(dv.data_ptr(), 1000, 0, 1000, idx.data_ptr(), dist.data_ptr(), rsum.data_ptr(),
                                   stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    dm = d.copy()
    np.fill_diagonal(dm, np.inf)
    assert np.array_equal(idx.cpu().numpy(), dm.argmin(1))
    assert np.array_equal(dist.cpu().numpy().astype(np.float64), dm.min(1))
    assert np.allclose(rsum.cpu().numpy(), d.sum(1), rtol=1e-5)  # float partial sums per column tile


# ---------------------------------------------------------------- device-resident entry points
def test_device_resident_batch_equals_host_batch(engine):
    import torch
    songs = [song_f32(200 + i, s) for i, s in enumerate([3.0, 5.2, 2.0])]
    host = engine.analyze_f32(songs)
    offs, total = [], 0
    for x in songs:
        offs.append(total)
        total += (len(x) + 63) // 64 * 64 + 64
    buf = torch.zeros(total, dtype=torch.float32, device="cuda")
    for o, x in zip(offs, songs):
        buf[o:o + len(x)] = torch.from_numpy(x).cuda()
    out = torch.zeros(len(songs) * 8, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    engine.analyze_device(bliss_b200.FMT_F32, buf.data_ptr(), offs, [len(x) for x in songs], out.data_ptr(), stream=st)
    torch.cuda.synchronize()
    got = np.frombuffer(out.cpu().numpy().tobytes(), dtype=bliss_b200.RESULT_DTYPE)
    assert got.tobytes() == host.tobytes()
    # the spectral-only kernel (BASELINE.json configs[1]) gives the same frequency rating up to float summation
    # order: it cuts a song into smaller parts (and is a separate instantiation of the kernel) ...
    freq = torch.zeros(len(songs), dtype=torch.float32, device="cuda")
    engine.spectral_device(bliss_b200.FMT_F32, buf.data_ptr(), offs, [len(x) for x in songs], freq.data_ptr(), stream=st)
    torch.cuda.synchronize()
    f1 = freq.cpu().numpy()
    assert np.max(np.abs(f1 - host["frequency"]) / np.abs(host["frequency"])) <= 1e-5
    # ... and is itself deterministic and independent of the batch
    freq2 = torch.zeros(1, dtype=torch.float32, device="cuda")
    engine.spectral_device(bliss_b200.FMT_F32, buf.data_ptr(), offs[1:2], [len(songs[1])], freq2.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert freq2.cpu().numpy()[0] == f1[1]


def test_parallel_nearest_neighbours_world1(engine, oracle):
    """bliss_b200.parallel on one rank: nearest neighbour of every song == argmin over bl_distance."""
    import torch
    from bliss_b200 import parallel
    rng = np.random.default_rng(21)
    v = (rng.standard_normal((3000, 4)) * np.array([5, 3, 8, 6])).astype(np.float32)
    v[100] = v[2000]  # an exact collision: distance 0, lowest index wins for third parties
    idx, dst = parallel.nearest_neighbours(engine, torch.from_numpy(v).cuda())
    torch.cuda.synchronize()
    d = oracle.distance_matrix(v)
    np.fill_diagonal(d, np.inf)
    assert np.array_equal(idx.cpu().numpy(), d.argmin(1))
    assert np.array_equal(dst.cpu().numpy(), d.min(1))
    slab = parallel.distance_slab(engine, torch.from_numpy(v[:257]).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(slab.cpu().numpy(), oracle.distance_matrix(v[:257]))


def test_empty_and_tiny_batches(engine):
    assert len(engine.analyze_s16([], [])) == 0
    assert len(engine.analyze_f32([])) == 0
    res = engine.analyze_s16([np.zeros(0, dtype=np.int16), song_s16(3, 2.0)], [0, 2])
    assert res[0]["status"] & bliss_b200.engine.SONG_TOO_SHORT
    assert res[1]["status"] == 0
    assert engine.distance_matrix(np.zeros((0, 4), dtype=np.float32)).shape == (0, 0)
    one = engine.distance_matrix(np.array([[1, 2, 3, 4]], dtype=np.float32))
    assert one.shape == (1, 1) and one[0, 0] == 0


def test_mono_song_full_analysis(engine, oracle):
    """channels == 1 through every analyser (reachable by direct calls, reference include/bliss.h:184-217)."""
    pcm = song_s16(31, 6.0)[::2].copy()
    res = engine.analyze_s16([pcm], [6], channels=[1])
    ref = oracle.analyze(pcm, 6, channels=1)
    check_song(res[0], ref, tag="mono")


def test_long_song_envelope_and_counts(engine, oracle):
    """A 3-minute song (the benchmark length): exact onset count and hop energies."""
    pcm = song_s16(55, 180.0, decorrelate=True, gain=0.7)
    res = engine.analyze_s16([pcm], [180])
    ref = oracle.analyze(pcm, 180)
    check_song(res[0], ref, tag="3min")
    E, Eo = engine.envelope_energy(pcm), oracle.envelope_energy(pcm)
    assert np.max(np.abs(E - Eo) / np.maximum(Eo, 1e-300)) <= 2.4e-7 and np.mean(E == Eo) > 0.99


# ---------------------------------------------------------------- envelope accumulation chain: edge cases
def _tonal_song(seed, seconds, freq, noise):
    """A strong sinusoid over very quiet noise: the hop spectra put nearly all their energy into one
    bin, so the float accumulation jumps several binades at once (multi-binade crossings)."""
    rng = np.random.default_rng(seed)
    n = int(22050 * seconds)
    t = np.arange(n) / 22050.0
    mono = 12000.0 * np.sin(2 * np.pi * freq * t) * (0.6 + 0.4 * np.sin(2 * np.pi * 1.7 * t)) + noise * rng.standard_normal(n)
    pcm = np.empty(2 * n, dtype=np.int16)
    pcm[0::2] = np.clip(np.rint(mono), -32768, 32767)
    pcm[1::2] = np.clip(np.rint(mono * 0.93), -32768, 32767)
    return pcm


def test_envelope_chain_tonal_and_quiet_songs(engine, oracle):
    cases = [_tonal_song(1, 7.0, 3000.0, 2.0), _tonal_song(2, 5.0, 9000.0, 0.6), _tonal_song(3, 6.0, 150.0, 30.0),
             (song_s16(71, 6.0, decorrelate=True).astype(np.int32) // 300).astype(np.int16)]  # a very quiet song
    for i, pcm in enumerate(cases):
        E, Eo = engine.envelope_energy(pcm), oracle.envelope_energy(pcm)
        assert np.max(np.abs(E - Eo) / np.maximum(Eo, 1e-300)) <= 2.4e-7, i
        assert np.mean(E == Eo) > 0.98, (i, float(np.mean(E == Eo)))
        res = engine.analyze_s16([pcm], [max(1, len(pcm) // 44100)])
        check_song(res[0], oracle.analyze(pcm, max(1, len(pcm) // 44100)), tag=f"tonal {i}")


def test_envelope_fast_chain_equals_slow_chain(engine):
    """The predicted-binade accumulation and its fallback (binade-by-binade scan) are the same function:
    forcing the fallback for every hop changes no bit of E[m], for native int16 and for float32 input."""
    songs = [song_s16(81, 12.0, decorrelate=True), _tonal_song(4, 4.0, 5000.0, 1.0), song_s16(82, 3.1, gain=0.05)]
    f32 = [song_f32(83, 6.0), song_f32(84, 3.3)]
    fast = [engine.envelope_energy(p) for p in songs]
    fast_rec = engine.analyze_f32(f32)
    engine.debug_flags(bliss_b200.engine.DEBUG_SLOW_CHAIN)
    try:
        slow = [engine.envelope_energy(p) for p in songs]
        slow_rec = engine.analyze_f32(f32)
    finally:
        engine.debug_flags(0)
    for a, b in zip(fast, slow):
        assert np.array_equal(a, b)
    assert fast_rec.tobytes() == slow_rec.tobytes()


def test_envelope_hop_counts_around_warp_and_cta_boundaries(engine, oracle):
    """A warp owns a run of consecutive hops (a power of two, 32..128) and a CTA eight of them (n_hops = 2 F - 2
    is always even): songs whose hop count ends just before / on / after those boundaries."""
    for n_hops in (18, 30, 32, 34, 62, 64, 66, 126, 128, 130, 254, 256, 258, 510, 512, 514, 1022, 1024, 1026, 1090):
        F = (n_hops + 2) // 2
        pcm = np.resize(song_s16(500 + n_hops, 1.0, decorrelate=True), 512 * F + 37)
        E, Eo = engine.envelope_energy(pcm), oracle.envelope_energy(pcm)
        assert E.shape == Eo.shape == (2 * F,)
        assert np.max(np.abs(E - Eo) / np.maximum(Eo, 1e-300)) <= 2.4e-7, n_hops
        assert np.all(E[:n_hops] > 0) and np.all(E[n_hops:] == 0), n_hops


def test_f32_song_is_independent_of_its_neighbours(engine):
    """The tensor copies of pass 1 also bring in rows of the neighbouring songs (FIR halo); they must read
    as zeros: a float32 song gives the same record alone, first, last or in the middle of a batch."""
    x = song_f32(91, 5.0)
    a, b = song_f32(92, 2.0), (song_f32(93, 3.0) * 3.0).astype(np.float32)
    alone = engine.analyze_f32([x])[0].tobytes()
    assert engine.analyze_f32([x, a, b])[0].tobytes() == alone
    assert engine.analyze_f32([a, x, b])[1].tobytes() == alone
    assert engine.analyze_f32([b, a, x])[2].tobytes() == alone


def test_song_with_silent_passages(engine, oracle):
    """Digital silence inside a song: hop energies of exactly zero (the accumulation never becomes a normal
    float) next to ordinary hops."""
    pcm = song_s16(95, 8.0, decorrelate=True)
    pcm[2 * 22050 * 2:2 * 22050 * 4] = 0  # seconds 2..4 silent
    pcm[:5001] = 0   # the reference's amplitude analyser skips the zeros in front of the first ...
    pcm[-3000:] = 0  # ... and behind the last non-zero sample (reference src/amplitude_sort.c:26-31)
    E, Eo = engine.envelope_energy(pcm), oracle.envelope_energy(pcm)
    assert np.array_equal(E == 0, Eo == 0)
    assert np.max(np.abs(E - Eo) / np.maximum(Eo, 1e-300)) <= 2.4e-7
    res = engine.analyze_s16([pcm], [8])
    check_song(res[0], oracle.analyze(pcm, 8), tag="silent passages")


def test_leading_and_trailing_silence_f32(engine, oracle):
    """Float32 input that starts and ends with digital silence (the trimmed zeros must not be counted in the
    amplitude histogram), plus a song whose only sound is in the middle."""
    x = song_f32(96, 6.0)
    x[:7001] = 0.0
    x[-9000:] = 0.0
    y = np.zeros(5 * 44100, dtype=np.float32)
    y[60000:150000] = song_f32(97, 5.0)[60000:150000]
    res = engine.analyze_f32([x, y])
    for i, z in enumerate([x, y]):
        pcm = oracle.frontend_f32(z)
        ref = oracle.analyze(pcm, len(z) // 44100)
        check_song(res[i], ref, tag=f"silence f32 {i}")
        assert float(res[i]["amplitude"]) == ref["amplitude"], i


def test_large_ragged_batch_properties(engine, oracle):
    """A few hundred float32 songs of mixed lengths in one call (several pass-1 parts, envelope CTAs and warps
    per song; duplicates; shuffled order): every record depends on its song alone - duplicates are
    byte-identical, a permutation of the batch permutes the records - and a sample of them matches the oracle."""
    rng = np.random.default_rng(2024)
    base = [song_f32(700 + i, float(s)) for i, s in enumerate(rng.uniform(1.2, 24.0, size=48))]
    base.append(song_f32(760, 95.0))  # several envelope CTAs and pass-1 parts
    order = rng.integers(0, len(base), size=320)
    batch = [base[i] for i in order]
    res = engine.analyze_f32(batch)
    first = {}
    for pos, i in enumerate(order):
        if i in first:
            assert res[pos].tobytes() == res[first[i]].tobytes(), (pos, i)
        else:
            first[i] = pos
    perm = rng.permutation(len(batch))
    res2 = engine.analyze_f32([batch[j] for j in perm])
    assert res2.tobytes() == res[perm].tobytes()
    for i in (0, 7, 19, 48):
        x = base[i]
        check_song(res[first[i]], oracle.analyze(oracle.frontend_f32(x), len(x) // 44100), tag=f"large batch song {i}")


def test_distance_nearest_column_splits(engine, oracle):
    """A slab of few rows against many columns takes the column-split path (gridDim.y > 1, atomicMin merge):
    same winner as the unsplit scan, including the lowest-index rule for ties across split boundaries."""
    import torch
    rng = np.random.default_rng(33)
    n = 40000
    v = (rng.standard_normal((n, 4)) * np.array([7, 5, 9, 6])).astype(np.float32)
    v[123] = v[37000]      # exact collisions far apart (different column ranges)
    v[30001] = v[5]
    v[5000] = v[123]       # three-way tie with 37000: row 123 must report the lower index (5000)
    dv = torch.from_numpy(v).cuda()
    st = torch.cuda.current_stream().cuda_stream
    rows0, nr = 100, 300   # rows 100..399 only: one row block -> many splits
    idx = torch.empty(nr, dtype=torch.int32, device="cuda")
    dist = torch.empty(nr, dtype=torch.float32, device="cuda")
    engine.distance_nearest_device(dv.data_ptr(), n, rows0, nr, idx.data_ptr(), dist.data_ptr(), 0, stream=st)
    idx2 = torch.empty(nr, dtype=torch.int32, device="cuda")
    dist2 = torch.empty(nr, dtype=torch.float32, device="cuda")
    rsum = torch.empty(nr, dtype=torch.float64, device="cuda")  # asking for row sums forces the unsplit kernel
    engine.distance_nearest_device(dv.data_ptr(), n, rows0, nr, idx2.data_ptr(), dist2.data_ptr(), rsum.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(idx.cpu().numpy(), idx2.cpu().numpy()) and np.array_equal(dist.cpu().numpy(), dist2.cpu().numpy())
    got_i, got_d = idx.cpu().numpy(), dist.cpu().numpy()
    assert got_i[123 - rows0] == 5000 and got_d[123 - rows0] == 0.0
    for r in (100, 123, 250, 399):
        d = np.array([oracle.distance(v[r], v[j]) for j in range(n)], dtype=np.float32)
        d[r] = np.inf
        assert got_i[r - rows0] == int(np.argmin(d)) and got_d[r - rows0] == d.min(), r


# ---------------------------------------------------------------- stage-level parity (one kernel at a time)
def test_stage_frequency_spectrum_per_bin(engine, oracle):
    """pass 1 alone: the per-bin power sum_f |X_d|^2 (before the dB compression of the rating hides per-bin
    errors) against the oracle's double FFT rounded to float (reference src/frequency_sort.c:83-93), for
    stereo (decorrelated channels), mono, and a float32 song after the front-end."""
    cases = [(song_s16(61, 12.0, decorrelate=True), 2), (song_s16(62, 4.0, gain=0.2), 2), (song_s16(63, 7.0)[::2].copy(), 1),
             (oracle.frontend_f32(song_f32(64, 9.0)), 2), (_tonal_song(5, 5.0, 2500.0, 3.0), 2)]
    worst = 0.0
    for i, (pcm, ch) in enumerate(cases):
        ps, ref = engine.frequency_spectrum(pcm, ch), oracle.frequency_spectrum(pcm, ch)
        assert ps[0] == 0 and ps[256] == 0
        err = np.abs(ps[1:256].astype(np.float64) - ref[1:256]) / (ref[1:256] + 1e-7 * ref[1:256].max())
        worst = max(worst, float(err.max()))
        assert err.max() <= 2e-5, (i, int(err.argmax()) + 1, float(err.max()))
        # the scalar epilogue of the rating is the oracle's on the same spectrum
        f_gpu = float(engine.analyze_s16([pcm], [max(1, len(pcm) // (22050 * ch))], channels=[ch], what=bliss_b200.DO_FREQUENCY)[0]["frequency"])
        assert rel(f_gpu, oracle.frequency_from_spectrum(ps)) <= 1e-6, i
    print(f"per-bin spectrum: worst relative error {worst:.3g}")


def test_stage_histogram_and_bounds(engine):
    """pass 1's integer side products: the exact histogram of the 3 807 relevant sample values and the
    first / last non-zero sample (reference src/amplitude_sort.c:26-39), against numpy."""
    pcm = song_s16(65, 6.0, decorrelate=True, gain=0.03)
    pcm[:1234] = 0
    pcm[-77:] = 0
    h, first, last = engine.histogram(pcm)
    want = np.bincount(pcm.astype(np.int64) + 32768, minlength=65536)[30864:34671]
    nz = np.flatnonzero(pcm)
    assert np.array_equal(h, want) and first == nz[0] and last == nz[-1]
    loud = song_s16(66, 3.0, gain=1.9)  # most samples outside the window: predicated atomics
    h2, _, _ = engine.histogram(loud)
    assert np.array_equal(h2, np.bincount(loud.astype(np.int64) + 32768, minlength=65536)[30864:34671])


def test_stage_envelope_tail_on_oracle_energies(engine, oracle):
    """logcomp + tail kernels alone, fed the ORACLE's hop energies: onset count and tempo exact, attack to
    float rounding (device log() vs glibc's), against orc_envelope_tail (reference
    src/tempo_atk_sort.c:184-287) - independent of the envelope kernel's FFT."""
    for i, (pcm, dur) in enumerate([(song_s16(67, 14.0, decorrelate=True), 14), (song_s16(68, 2.0), 2),
                                    (_tonal_song(6, 8.0, 700.0, 5.0), 8), (song_s16(69, 61.0, gain=0.3), 61)]):
        E = oracle.envelope_energy(pcm)
        ref = oracle.envelope_tail(E, len(pcm), dur)
        got = engine.envelope_tail(E, len(pcm), dur)
        assert got["beat"] == ref["beat"], (i, got, ref)
        assert got["tempo"] == ref["tempo"], (i, got, ref)
        assert rel(got["attack"], ref["attack"]) <= 1e-6, (i, got, ref)
    # a synthetic envelope with exact ties and plateaus (differences of exactly 0 and exactly epsilon-scale)
    E = np.zeros(400)
    E[10:390:20] = 3.0
    E[15:390:20] = 3.0
    E[-2:] = 0
    ref = oracle.envelope_tail(E, 200 * 512 + 5, 4)
    got = engine.envelope_tail(E, 200 * 512 + 5, 4)
    assert got["beat"] == ref["beat"] and got["tempo"] == ref["tempo"] and rel(got["attack"], ref["attack"]) <= 1e-6


def test_bl_distance_file_two_different_files(engine, oracle, tmp_path):
    """reference src/analyze.c:105-125 on two DIFFERENT files: the distance equals bl_distance of the two
    analysed vectors bit for bit, and of the oracle's vectors for the same PCM to 1e-4."""
    import wave
    L = bliss_b200.load()
    pcm = song_s16(70, 9.0, decorrelate=True, gain=0.6)
    wav = str(tmp_path / "other.wav")
    with wave.open(wav, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(22050)
        w.writeframes(pcm.tobytes())
    flac = os.path.join(GOLDEN_DIR, "song.flac").encode()
    s1, s2 = bliss_b200.BlSong(), bliss_b200.BlSong()
    d = L.bl_distance_file(flac, wav.encode(), ctypes.byref(s1), ctypes.byref(s2))
    v1 = [s1.force_vector.tempo, s1.force_vector.amplitude, s1.force_vector.frequency, s1.force_vector.attack]
    v2 = [s2.force_vector.tempo, s2.force_vector.amplitude, s2.force_vector.frequency, s2.force_vector.attack]
    assert d > 1.0 and d == np.float32(oracle.distance(v1, v2))
    assert d == L.bl_distance(s1.force_vector, s2.force_vector)
    c = L.bl_cosine_similarity_file(flac, wav.encode(), ctypes.byref(s1), ctypes.byref(s2))
    assert c == np.float32(oracle.cosine_similarity(v1, v2)) and -1.0 <= c < 1.0
    ref2 = oracle.analyze(pcm, 9)
    from test_oracle import GOLDEN_S16
    o1 = [GOLDEN_S16[k] for k in ("tempo", "amplitude", "frequency", "attack")]
    o2 = [ref2[k] for k in ("tempo", "amplitude", "frequency", "attack")]
    assert rel(d, oracle.distance(o1, o2)) <= 1e-4
    assert s2.nSamples == len(pcm) and s2.duration == 9
    L.bl_free_song(ctypes.byref(s1)); L.bl_free_song(ctypes.byref(s2))


def test_sub_batches_and_pipelined_calls_are_byte_identical(engine):
    """The device-resident path cuts a batch into sub-batches (tail of one under the envelope kernel of the next, two
    streams) and a job may enqueue several batches before joining: records are byte-identical to one sub-batch per call."""
    import torch
    songs = [song_f32(800 + i, s) for i, s in enumerate([3.0, 7.5, 2.2, 12.0, 4.0, 2.0, 9.1, 5.5, 3.3, 6.0, 2.7])]
    offs, total = [], 0
    for x in songs:
        offs.append(total)
        total += (len(x) + 63) // 64 * 64 + 64
    buf = torch.zeros(total, dtype=torch.float32, device="cuda")
    for o, x in zip(offs, songs):
        buf[o:o + len(x)] = torch.from_numpy(x).cuda()
    lens = [len(x) for x in songs]
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        ref = torch.zeros(len(songs) * 8, dtype=torch.int32, device="cuda")
        engine.analyze_device(bliss_b200.FMT_F32, buf.data_ptr(), offs, lens, ref.data_ptr(), stream=st.cuda_stream)
        st.synchronize()
        want = ref.cpu().numpy().tobytes()
        small = bliss_b200.Engine(0)
        try:
            small.configure_sub_batch(3)  # 3 + 3 + 3 + 2 songs: all four slots in rotation
            out = torch.zeros_like(ref)
            small.analyze_device(bliss_b200.FMT_F32, buf.data_ptr(), offs, lens, out.data_ptr(), stream=st.cuda_stream)
            st.synchronize()
            assert out.cpu().numpy().tobytes() == want
            # three pipelined calls (5 + 4 + 2 songs) into one result array, one join
            out2 = torch.zeros_like(ref)
            for a, b in ((0, 5), (5, 9), (9, 11)):
                small.analyze_device(bliss_b200.FMT_F32, buf.data_ptr(), offs[a:b], lens[a:b], out2.data_ptr() + 32 * a,
                                     stream=st.cuda_stream, wait=False)
            small.join(st.cuda_stream)
            # work enqueued on the caller's stream after the join sees the results
            copy = out2.clone()
            st.synchronize()
            assert copy.cpu().numpy().tobytes() == want
        finally:
            small.close()


def test_bl_analyze_from_many_threads(oracle):
    """The drop-in wrappers take an engine out of a pool per call: eight threads analysing at once (the reference is not
    thread-safe at all, reference src/tempo_atk_sort.c:94,295) all get the single-threaded record, byte for byte."""
    import threading
    L = bliss_b200.load()
    flac = os.path.join(GOLDEN_DIR, "song.flac").encode()

    def analyse():
        s = bliss_b200.BlSong()
        rc = L.bl_analyze(flac, ctypes.byref(s))
        rec = (rc, s.force, s.force_vector.tempo, s.force_vector.amplitude, s.force_vector.frequency, s.force_vector.attack,
               s.nSamples, s.calm_or_loud)
        env = bliss_b200.EnvelopeResult()
        L.bl_envelope_sort(ctypes.byref(s), ctypes.byref(env))
        rec += (env.tempo, env.attack, L.bl_amplitude_sort(ctypes.byref(s)), L.bl_frequency_sort(ctypes.byref(s)))
        L.bl_free_song(ctypes.byref(s))
        return rec

    want = analyse()
    assert want[0] == 1 and want[8:] == (want[2], want[5], want[3], want[4])
    got, errs = [], []

    def worker():
        try:
            for _ in range(3):
                got.append(analyse())
        except Exception as e:  # pragma: no cover
            errs.append(e)

    threads = [threading.Thread(target=worker) for _ in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs and len(got) == 24 and all(g == want for g in got)


def test_default_stream_is_ordered_with_the_callers_work(engine):
    """torch.cuda.current_stream().cuda_stream is 0 on the legacy default stream; the binding must run the kernel ON that
    stream (cudaStreamLegacy), not on the engine's own: a slow producer in front and a consumer behind see ordinary
    stream order without any synchronize in between."""
    import torch
    from bliss_b200 import parallel
    rng = np.random.default_rng(5)
    v = (rng.standard_normal((4096, 4)) * 9).astype(np.float32)
    want_idx, want_dst = parallel.nearest_neighbours(engine, torch.from_numpy(v).cuda())
    torch.cuda.synchronize()
    want_idx, want_dst = want_idx.cpu().numpy(), want_dst.cpu().numpy()
    assert torch.cuda.current_stream().cuda_stream == 0
    a = torch.randn(6144, 6144, device="cuda")
    staged = torch.from_numpy(v).cuda()
    torch.cuda.synchronize()
    for _ in range(3):
        table = torch.zeros(4096, 4, device="cuda")
        b = a
        for _ in range(6):
            b = b @ a  # tens of milliseconds of work queued in front of the copy below
        table.copy_(staged)  # the real vectors arrive only after the matmuls
        idx, dst = parallel.nearest_neighbours(engine, table)
        got_idx, got_dst = idx.cpu().numpy(), dst.cpu().numpy()  # .cpu() waits on the default stream only
        assert np.array_equal(got_idx, want_idx) and np.array_equal(got_dst, want_dst)
        del b


def test_cosine_nearest_matches_cosine_matrix(engine, oracle):
    """Fused cosine epilogue: per row the most similar other song = argmax over the (bit-exact) cosine matrix, lowest index on
    ties, including exact duplicates, scaled copies (similarity rounds to 1.0 for several columns) and a zero vector (NaN)."""
    import torch
    rng = np.random.default_rng(44)
    v = (rng.standard_normal((2500, 4)) * np.array([6, 4, 9, 7])).astype(np.float32)
    v[100] = v[2000]
    v[300] = v[2000] * 2.0       # same direction, different length
    v[301] = v[2000] * 0.5
    v[777] = 0.0                 # |v| = 0: every similarity with it is NaN and never wins
    c = engine.distance_matrix(v, cosine=True)
    assert c[3, 9] == np.float32(oracle.cosine_similarity(v[3], v[9]))
    dv = torch.from_numpy(v).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for row0, nr in ((0, 2500), (1990, 300)):
        idx = torch.full((nr,), -7, dtype=torch.int32, device="cuda")
        sim = torch.zeros(nr, dtype=torch.float32, device="cuda")
        engine.cosine_nearest_device(dv.data_ptr(), 2500, row0, nr, idx.data_ptr(), sim.data_ptr(), stream=st)
        gi, gs = idx.cpu().numpy(), sim.cpu().numpy()
        for r in range(nr):
            i = row0 + r
            if i == 777:
                assert gi[r] == -1
                continue
            rowv = c[i].astype(np.float64).copy()
            rowv[i] = -np.inf
            rowv[np.isnan(rowv)] = -np.inf
            assert gs[r] == np.float32(rowv.max()) and gi[r] == int(np.argmax(rowv)), (i, gi[r], gs[r], int(np.argmax(rowv)), rowv.max())
    from bliss_b200 import compat as bliss
    order, vals = bliss.playlist(engine, v, 2000, k=6, metric="cosine")
    assert set(order[:4].tolist()) == {100, 300, 301, 2000} and np.all(vals[:4] >= np.float32(0.9999999))


def test_f32_exact_path_equals_resampler_plus_native_analysis(engine, oracle):
    """blx_analyze_batch_f32_exact: float32 songs through the libswresample-exact resampler on the device, then the native
    int16 pipeline - byte-identical to analysing the oracle-resampled PCM, for mono 44.1 kHz, stereo 48 kHz and a ragged batch."""
    mono = [song_f32(900 + i, s) for i, s in enumerate([3.0, 7.3, 2.0])]
    got = engine.analyze_f32_exact(mono, in_rate=44100, channels=1)
    for i, x in enumerate(mono):
        pcm = oracle.resample_to_s16(x.view(np.int32), oracle.RS_F32, 32, 1, 44100)
        want = engine.analyze_s16([pcm], [len(x) // 44100])[0]
        assert got[i].tobytes() == want.tobytes(), i
        check_song(got[i], oracle.analyze(pcm, len(x) // 44100), tag=f"f32 exact mono {i}")
    a, b = song_f32(910, 5.0, rate=48000), song_f32(911, 5.0, rate=48000)
    st = np.stack([a, 0.7 * b], axis=1).reshape(-1).astype(np.float32)
    got2 = engine.analyze_f32_exact([st], in_rate=48000, channels=2)[0]
    pcm2 = oracle.resample_to_s16(st.view(np.int32), oracle.RS_F32, 32, 2, 48000)
    assert got2.tobytes() == engine.analyze_s16([pcm2], [5])[0].tobytes()


def test_twenty_minute_song(engine, oracle):
    """Maximum sizes: a 20-minute native int16 stereo song (52.9 M samples, 206 k envelope hops, 404 envelope CTAs, 101 pass-1
    parts) against the oracle - onset count exact, the amplitude histogram beyond 2^24 counts per bin where the reference's float
    counters stop (reference src/amplitude_sort.c:33-39) - and the same song after a 44.1 kHz float round trip through the
    exact resampler path."""
    base = song_s16(123, 60.0, decorrelate=True, gain=0.08)  # quiet: the central histogram bins get far more than 2^24 hits
    pcm = np.tile(base, 20)
    pcm[1::7] //= 3
    res = engine.analyze_s16([pcm], [1200])[0]
    ref = oracle.analyze(pcm, 1200)
    check_song(res, ref, tag="20 min")
    assert float(res["amplitude"]) == ref["amplitude"]
    E, Eo = engine.envelope_energy(pcm), oracle.envelope_energy(pcm)
    assert E.shape == Eo.shape and np.array_equal(E, Eo)


def test_extreme_waveforms(engine, oracle):
    """Signals at the edges of the int16 range and of the filters' pass bands: clipped full-scale square wave, Nyquist
    alternation (over faint noise: with nothing else in the signal the band levels ARE the rounding noise of whichever
    float FFT computed them, the reference's included), a sparse impulse train, a large DC offset under faint noise,
    anti-phase stereo (L = -R: the frequency analyser sees an all-zero mono mix and returns NaN, as the reference does), and white noise at full scale.
    Amplitude is bit-exact, the onset count exact, E[m] within float rounding of the reference's on every hop."""
    rng = np.random.default_rng(20251017)
    n = 22050 * 8
    t = np.arange(n)
    sq = np.where((t // 50) % 2 == 0, 32767, -32768).astype(np.int16)
    nyq = (np.where(t % 2 == 0, 30000, -30000) + rng.integers(-200, 201, n)).astype(np.int16)
    imp = np.zeros(n, dtype=np.int16)
    imp[::4410] = 32767
    imp[2205::4410] = -32768
    dc = (20000 + rng.integers(-3, 4, n)).astype(np.int16)
    base = song_s16(91, 8.0)[0::2]
    white = rng.integers(-32768, 32768, n).astype(np.int16)

    def stereo(left, right=None):
        return np.stack([left, left if right is None else right], axis=1).reshape(-1)

    anti = stereo(base, (-base.astype(np.int32)).clip(-32768, 32767).astype(np.int16))
    cases = {"square": stereo(sq), "nyquist": stereo(nyq), "impulses": stereo(imp), "dc": stereo(dc), "antiphase": anti,
             "white": stereo(white, np.roll(white, 7))}
    problems = []
    for name, pcm in cases.items():
        ref = oracle.analyze(pcm, 8)
        res = engine.analyze_s16([pcm], [8])[0]
        E, Eo = engine.envelope_energy(pcm), oracle.envelope_energy(pcm)
        try:
            assert float(res["amplitude"]) == np.float32(ref["amplitude"]), "amplitude"
            check_song(res, ref, tag=name)
            assert np.max(np.abs(E - Eo) / np.maximum(Eo, 1e-300)) <= 2.4e-7, "E[m]"
        except AssertionError as e:
            problems.append((name, str(e)[:300]))
    assert not problems, problems
