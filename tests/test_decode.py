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


def test_decode_error_paths(tmp_path, capfd):
    L, s, rc = decode(tmp_path / "missing.flac")
    assert rc == -2  # BL_UNEXPECTED
    L.bl_free_song(ctypes.byref(s))  # a failed decode can still be freed (reference examples/analyze.c:15-17,50-52)
    (tmp_path / "junk.wav").write_bytes(b"not audio at all" * 10)
    assert decode(tmp_path / "junk.wav")[2] == -2
    write_wav(tmp_path / "r48.wav", song_s16(6, 1.0), 48000, 2, 1, 16)  # needs a resampler this build does not carry
    assert decode(tmp_path / "r48.wav")[2] == -2
    assert "resampler" in capfd.readouterr().err
    assert L.bl_analyze(str(tmp_path / "missing.flac").encode(), ctypes.byref(bliss_b200.BlSong())) == -2


@pytest.mark.gpu
def test_wav_44k_routes_through_the_gpu_frontend(tmp_path, engine, oracle):
    x = song_f32(7, 4.0)
    xs = np.stack([x * 0.9, x * 0.5], axis=1)  # stereo: the host mixes to mono before the 2:1 front-end
    write_wav(tmp_path / "f32m.wav", x, 44100, 1, 3, 32)
    write_wav(tmp_path / "s16s.wav", np.round(xs * 32767).astype(np.int16), 44100, 2, 1, 16)
    write_wav(tmp_path / "s24m.wav", np.round(x * 8388607).astype(np.int32), 44100, 1, 1, 24)
    L = bliss_b200.load()
    # float32 mono: exactly the engine's float path
    s = bliss_b200.BlSong()
    assert L.bl_analyze(str(tmp_path / "f32m.wav").encode(), ctypes.byref(s)) in (0, 1, 2)
    ref = engine.analyze_f32([x])[0]
    assert s.resampled == 1 and s.nSamples == 2 * (len(x) // 2) and s.duration == 4
    assert np.array_equal(pcm_of(s), oracle.frontend_f32(x))
    assert (s.force_vector.tempo, s.force_vector.amplitude, s.force_vector.frequency, s.force_vector.attack, s.force) == (
        ref["tempo"], ref["amplitude"], ref["frequency"], ref["attack"], ref["force"])
    L.bl_free_song(ctypes.byref(s))
    # int16 stereo and packed int24 mono: scaled to [-1, 1), mixed, then the same front-end
    for name, mono in (("s16s.wav", ((np.round(xs * 32767).astype(np.int16).astype(np.float32) / 32768.0).sum(axis=1) / 2).astype(np.float32)),
                       ("s24m.wav", (np.round(x * 8388607).astype(np.int32).astype(np.float32) / 8388608.0).astype(np.float32))):
        s = bliss_b200.BlSong()
        assert L.bl_audio_decode(str(tmp_path / name).encode(), ctypes.byref(s)) == 0, name
        assert np.array_equal(pcm_of(s), oracle.frontend_f32(mono)), name
        L.bl_free_song(ctypes.byref(s))
