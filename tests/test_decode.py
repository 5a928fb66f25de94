"""bl_audio_decode / bl_analyze on files (row N1): RIFF/WAVE in the formats the host reader takes, the native
22 050 Hz int16 pass-through, the 44.1 kHz route through the GPU front-end, and the error paths. The FLAC route
is covered by the reference's fixture in test_oracle.py / test_gpu_parity.py."""
import ctypes
import os
import struct

import numpy as np
import pytest

import bliss_b200
from synth import song_f32, song_s16


def write_wav(path, data, rate, channels, fmt_tag, bits):
    """Minimal RIFF/WAVE writer: PCM (fmt_tag 1: int16 / packed int24 / int32) or IEEE float (fmt_tag 3)."""
    if bits == 24:
        a = np.asarray(data, dtype=np.int32)
        raw = np.stack([(a >> s) & 0xff for s in (0, 8, 16)], axis=-1).astype(np.uint8).tobytes()
    else:
        raw = np.ascontiguousarray(data).tobytes()
    block = channels * bits // 8
    hdr = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, fmt_tag, channels, rate, rate * block, block, bits) + b"data" + struct.pack("<I", len(raw))
    with open(path, "wb") as f:
        f.write(hdr + raw)


def decode(path):
    L = bliss_b200.load()
    s = bliss_b200.BlSong()
    rc = L.bl_audio_decode(str(path).encode(), ctypes.byref(s))
    return L, s, rc


def pcm_of(s):
    return np.ctypeslib.as_array(ctypes.cast(s.sample_array, ctypes.POINTER(ctypes.c_int16)), (s.nSamples,)).copy()


def test_wav_native_passthrough(tmp_path):
    pcm = song_s16(5, 2.5, decorrelate=True)
    write_wav(tmp_path / "a.wav", pcm, 22050, 2, 1, 16)
    L, s, rc = decode(tmp_path / "a.wav")
    assert rc == 0 and s.nSamples == len(pcm) and s.channels == 2 and s.sample_rate == 22050 and s.resampled == 0
    assert s.duration == 2 and s.nb_bytes_per_sample == 2
    assert np.array_equal(pcm_of(s), pcm)
    # reference src/decode.c:261-309: defaults for missing tags
    assert (s.title, s.artist, s.album, s.genre, s.tracknumber) == (b"<no title>", b"<no artist>", b"<no album>", b"<no genre>", b"")
    L.bl_free_song(ctypes.byref(s))
    assert s.sample_array is None and s.title is None


def test_wav_chunk_layouts(tmp_path):
    """16-bit PCM is read straight into the song's buffer (flac_reader.c: read_wav16_direct); the chunk walk in front of
    it has to cope with what real files carry: a LIST chunk of odd length (padded) before `data`, WAVE_FORMAT_EXTENSIBLE,
    a `data` length that overshoots the file (truncated download), a header pushed beyond the first 4 KB (whole-file path)."""
    pcm = song_s16(7, 1.5, decorrelate=True)
    raw = pcm.tobytes()
    fmt16 = struct.pack("<HHIIHH", 1, 2, 22050, 22050 * 4, 4, 16)
    fmt_ext = struct.pack("<HHIIHH", 0xFFFE, 2, 22050, 22050 * 4, 4, 16) + struct.pack("<HHI", 22, 16, 3) + struct.pack(
        "<H", 1) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"

    def riff(chunks):
        body = b"WAVE" + b"".join(k + struct.pack("<I", n if n is not None else len(d)) + d + (b"\x00" if len(d) & 1 else b"")
                                  for k, d, n in chunks)
        return b"RIFF" + struct.pack("<I", len(body)) + body

    cases = {
        "list.wav": (riff([(b"fmt ", fmt16, None), (b"LIST", b"INFOISFT\x05\x00\x00\x00abcd\x00", None), (b"data", raw, None)]), len(pcm)),
        "ext.wav": (riff([(b"fmt ", fmt_ext, None), (b"data", raw, None)]), len(pcm)),
        "short.wav": (riff([(b"fmt ", fmt16, None), (b"data", raw[:len(raw) // 2 + 2], len(raw))]), (len(raw) // 2 + 2) // 4 * 2),
        "late.wav": (riff([(b"fmt ", fmt16, None), (b"junk", bytes(6000), None), (b"data", raw, None)]), len(pcm)),
    }
    for name, (blob, n_expect) in cases.items():
        (tmp_path / name).write_bytes(blob)
        L, s, rc = decode(tmp_path / name)
        assert rc == 0, name
        got = pcm_of(s)
        assert len(got) == n_expect and np.array_equal(got, pcm[:len(got)]), name
        L.bl_free_song(ctypes.byref(s))
    (tmp_path / "empty.wav").write_bytes(riff([(b"fmt ", fmt16, None), (b"data", b"", None)]))
    assert decode(tmp_path / "empty.wav")[2] == -2


def test_decode_error_paths(tmp_path, capfd):
    L, s, rc = decode(tmp_path / "missing.flac")
    assert rc == -2  # BL_UNEXPECTED
    L.bl_free_song(ctypes.byref(s))  # a failed decode can still be freed (reference examples/analyze.c:15-17,50-52)
    (tmp_path / "junk.wav").write_bytes(b"not audio at all" * 10)
    assert decode(tmp_path / "junk.wav")[2] == -2
    six = np.zeros(6 * 500, dtype=np.int16)
    write_wav(tmp_path / "six.wav", six, 22050, 6, 1, 16)  # 5.1: no down-mix matrix in this build
    assert decode(tmp_path / "six.wav")[2] == -2
    assert "down-mix" in capfd.readouterr().err
    write_wav(tmp_path / "rate0.wav", np.zeros(2000, dtype=np.int16), 0, 2, 1, 16)  # a header that claims 0 Hz
    assert decode(tmp_path / "rate0.wav")[2] == -2
    assert "unusable stream parameters" in capfd.readouterr().err
    assert L.bl_analyze(str(tmp_path / "missing.flac").encode(), ctypes.byref(bliss_b200.BlSong())) == -2


@pytest.mark.gpu
def test_files_at_other_rates_and_formats_go_through_the_resampler(tmp_path, oracle):
    """bl_audio_decode on WAV files that are not int16 / 22 050 Hz: the GPU resampler (include/blx_resample.h) gives what the
    oracle restatement of libswresample gives, bit for bit - stereo stays stereo, mono is up-mixed, every bit depth."""
    x = song_f32(7, 3.0)
    xs = np.stack([x * 0.9, np.roll(x, 17) * 0.5], axis=1)
    L = bliss_b200.load()
    cases = [("f32m_44.wav", x, 44100, 1, 3, 32, oracle.RS_F32, x.view(np.int32)),
             ("s16s_44.wav", np.round(xs * 32767).astype(np.int16), 44100, 2, 1, 16, oracle.RS_S16, None),
             ("s24m_48.wav", np.round(x * 8388607).astype(np.int32), 48000, 1, 1, 24, oracle.RS_S32, None),
             ("s16s_48.wav", np.round(xs * 32767).astype(np.int16), 48000, 2, 1, 16, oracle.RS_S16, None),
             ("s32s_96.wav", np.round(xs * 2147483000).astype(np.int32), 96000, 2, 1, 32, oracle.RS_S32, None),
             ("u8s_8.wav", (np.round(xs[:20000] * 127) + 128).astype(np.uint8), 8000, 2, 1, 8, oracle.RS_U8, None),
             ("s24s_22.wav", np.round(xs * 8388607).astype(np.int32), 22050, 2, 1, 24, oracle.RS_S32, None)]
    for name, data, rate, ch, tag, bits, kind, raw in cases:
        write_wav(tmp_path / name, data, rate, ch, tag, bits)
        _, s, rc = decode(tmp_path / name)
        assert rc == 0, name
        if raw is None:
            raw = np.asarray(data).astype(np.int32) - (128 if bits == 8 else 0)
        want = oracle.resample_to_s16(raw.reshape(-1), kind, bits, ch, rate)
        got = pcm_of(s)
        assert s.resampled == 1 and s.channels == 2 and s.sample_rate == 22050 and s.nSamples == len(want), name
        assert s.duration == len(np.asarray(data).reshape(-1, ch)) // rate, name
        assert np.array_equal(got, want), (name, int(np.abs(got.astype(int) - want).max()))
        L.bl_free_song(ctypes.byref(s))


@pytest.mark.gpu
def test_cd_audio_flac_goes_through_the_resampler(tmp_path, oracle, monkeypatch):
    """The everyday input: 44.1 kHz / 16 bit / stereo FLAC (here from the test encoder; large enough for the threaded
    decoder). Decoded straight to int16, resampled on the GPU as int16, analysed: PCM equal to the oracle resampler's, force
    vector equal to the reference analysers' on that PCM. A 12-bit mono FLAC at 32 kHz takes the int32 route."""
    from flac_encode import encode
    x = song_f32(11, 6.0)
    xs = np.round(np.stack([x * 0.9, np.roll(x, 23) * 0.6], axis=1) * 32767).astype(np.int64)
    plan = lambda fi: dict(kind=["lpc", "fixed2", "fixed4"][fi % 3], stereo=[None, 8, 9, 10][fi % 4], lpc_order=1 + (fi * 3) % 12, porder=3)
    (tmp_path / "cd.flac").write_bytes(encode(xs, 16, 44100, 4096, plan, seed=3))
    want = oracle.resample_to_s16(xs.astype(np.int32).reshape(-1), oracle.RS_S16, 16, 2, 44100)
    for forced in (False, True):  # host decode + resampler call; then decode AND resampling fused on the device
        if forced:
            monkeypatch.setenv("BLX_FLAC_GPU_MIN_SAMPLES", "0")
        before = bliss_b200.load().blx_flac_accelerated_count()
        L, s, rc = decode(tmp_path / "cd.flac")
        assert L.blx_flac_accelerated_count() == before + (1 if forced else 0)
        assert rc == 0 and s.resampled == 1 and s.channels == 2 and s.sample_rate == 22050 and s.duration == 6
        assert s.nSamples == len(want) and np.array_equal(pcm_of(s), want), forced
        L.bl_free_song(ctypes.byref(s))
    s2 = bliss_b200.BlSong()
    assert L.bl_analyze(str(tmp_path / "cd.flac").encode(), ctypes.byref(s2)) in (0, 1)
    ref = oracle.analyze(want, 6)
    for k in ("tempo", "amplitude", "frequency", "attack"):
        assert abs(getattr(s2.force_vector, k) - ref[k]) <= 1e-4 * max(abs(ref[k]), 1e-30), k
    L.bl_free_song(ctypes.byref(s2))
    m = np.round(song_f32(12, 2.0)[:60000] * 2047).astype(np.int64).reshape(-1, 1)
    (tmp_path / "m12.flac").write_bytes(encode(m, 12, 32000, 1024, lambda fi: dict(kind="lpc", lpc_order=6), seed=4))
    L, s, rc = decode(tmp_path / "m12.flac")
    assert rc == 0 and s.resampled == 1
    assert np.array_equal(pcm_of(s), oracle.resample_to_s16(m.astype(np.int32).reshape(-1), oracle.RS_S16, 12, 1, 32000))
    L.bl_free_song(ctypes.byref(s))


@pytest.mark.gpu
def test_flac_frames_decode_on_the_device(tmp_path, monkeypatch):
    """Long FLAC streams are decoded on the GPU, one thread per frame (csrc/flacdec.cu): same samples as the host decoder
    for the reference's fixtures (forced onto the device path), a three-minute stream and encoder streams with every
    subframe type; a damaged stream is refused by the device and decoded by the host fall-back."""
    from flac_encode import encode
    from flac_util import read_pcm_file, repeat_flac
    L = bliss_b200.load()
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    files = [os.path.join(golden, n) for n in ("song.flac", "song_s32.flac", "song_s32_mono.flac")]
    (tmp_path / "long.flac").write_bytes(repeat_flac(files[0], 16))  # 177 s, above the default threshold
    files.append(str(tmp_path / "long.flac"))
    rng = np.random.default_rng(5)
    t = np.arange(1152 * 120)
    pcm = np.stack([8000 * np.sin(t / 37.0) + rng.standard_normal(len(t)) * 900, 6000 * np.sin(t / 41.0) + rng.standard_normal(len(t)) * 900],
                   axis=1).round().astype(np.int64)
    kinds = ["lpc", "fixed0", "fixed1", "fixed2", "fixed3", "fixed4", "verbatim"]
    plan = lambda fi: dict(kind=kinds[fi % 7], stereo=[None, 8, 9, 10][fi % 4], lpc_order=1 + (fi * 5) % 32, method=fi % 2,
                           porder=[0, 1, 3, 5][fi % 4], escape_parts=(0,) if fi % 6 == 5 else ())
    (tmp_path / "enc16.flac").write_bytes(encode(pcm, 16, 44100, 1152, plan, seed=1))
    (tmp_path / "enc24.flac").write_bytes(encode(pcm * 200, 24, 48000, 1152, plan, seed=2))
    files += [str(tmp_path / "enc16.flac"), str(tmp_path / "enc24.flac")]
    blob = bytearray(open(files[3], "rb").read())
    blob[len(blob) // 3] ^= 0x10
    (tmp_path / "damaged.flac").write_bytes(bytes(blob))
    files.append(str(tmp_path / "damaged.flac"))
    monkeypatch.delenv("BLX_FLAC_EMULATE", raising=False)
    for path in files:
        monkeypatch.setenv("BLX_FLAC_GPU", "0")
        ref = read_pcm_file(path)
        monkeypatch.setenv("BLX_FLAC_GPU", "1")
        if "long" not in path and "damaged" not in path:
            monkeypatch.setenv("BLX_FLAC_GPU_MIN_SAMPLES", "0")
        else:
            monkeypatch.delenv("BLX_FLAC_GPU_MIN_SAMPLES", raising=False)
        before = L.blx_flac_accelerated_count()
        got = read_pcm_file(path)
        assert L.blx_flac_accelerated_count() == before + (0 if "damaged" in path else 1), path
        assert got[1:] == ref[1:] and np.array_equal(got[0], ref[0]), path
    # and through the drop-in entry point: the golden force vector with the fixture decoded on the device
    monkeypatch.setenv("BLX_FLAC_GPU_MIN_SAMPLES", "0")
    s = bliss_b200.BlSong()
    assert L.bl_analyze(files[0].encode(), ctypes.byref(s)) == 1
    assert abs(s.force_vector.tempo - (-8.945454)) <= 1e-5 and abs(s.force_vector.attack - (-15.560563)) <= 1e-5
    L.bl_free_song(ctypes.byref(s))


def test_flac_frame_checksums_reject_damaged_frames(tmp_path):
    """A FLAC frame whose CRC-16 (or header CRC-8) does not match is dropped, not decoded as audio, and a sync code that
    happens to occur inside foreign bytes is not taken for a frame: the rest of the file decodes bit for bit."""
    from conftest import GOLDEN_DIR
    src = open(os.path.join(GOLDEN_DIR, "song.flac"), "rb").read()
    L, s, rc = decode(os.path.join(GOLDEN_DIR, "song.flac"))
    assert rc == 0
    good = pcm_of(s)
    L.bl_free_song(ctypes.byref(s))
    bad = bytearray(src)
    mid = len(bad) // 2
    for k in range(40):
        bad[mid + k] ^= 0x5A  # damage one frame's payload
    (tmp_path / "damaged.flac").write_bytes(bytes(bad))
    _, s2, rc2 = decode(tmp_path / "damaged.flac")
    assert rc2 == 0
    got = pcm_of(s2)
    L.bl_free_song(ctypes.byref(s2))
    lost = len(good) - len(got)
    assert 0 < lost <= 2 * 2 * 4096  # one (at most two) blocks of 4 096 stereo frames are gone, nothing else
    # everything in front of the damage is untouched, everything behind it follows after the gap
    k = next(i for i in range(0, len(got), 2) if got[i] != good[i] or got[i + 1] != good[i + 1])
    assert np.array_equal(got[:k], good[:k]) and np.array_equal(got[k:], good[k + lost:])
    # foreign bytes with a sync pattern in front of the first frame are skipped
    first = src.index(b"\xff\xf8", 4 + 38)
    (tmp_path / "junk.flac").write_bytes(src[:first] + b"\xff\xf8\xc9\x18" + bytes(range(64)) + src[first:])
    _, s3, rc3 = decode(tmp_path / "junk.flac")
    assert rc3 == 0 and np.array_equal(pcm_of(s3), good)
    L.bl_free_song(ctypes.byref(s3))
