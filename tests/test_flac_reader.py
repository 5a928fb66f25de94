"""The host FLAC decoder (bliss_b200/host/flac_reader.c) against streams built by tests/flac_encode.py: every subframe
type, LPC orders 1..32 (the specialised loops and the generic one), 32- and 64-bit synthesis, wasted bits, escaped and
5-bit Rice partitions, partition orders, all stereo modes, 8..24-bit samples, short last block. Decoding must give back
the PCM exactly. (The reference's own fixtures pin the decoder against libFLAC-encoded files: test_oracle.py.)"""
import numpy as np
import pytest

from flac_encode import encode
from flac_util import read_pcm_file


def _signal(rng, n, ch, bps):
    t = np.arange(n)
    amp = (1 << (bps - 1)) * 0.6
    x = np.stack([amp * (0.5 * np.sin(2 * np.pi * t / rng.uniform(20, 400)) + 0.2 * rng.standard_normal(n)) for _ in range(ch)], axis=1)
    if ch == 2:
        x[:, 1] = 0.7 * x[:, 0] + 0.3 * x[:, 1]  # correlated channels, as real stereo is
    return np.clip(np.round(x), -(1 << (bps - 1)), (1 << (bps - 1)) - 1).astype(np.int64)


KINDS = ["lpc", "fixed0", "fixed1", "fixed2", "fixed3", "fixed4", "verbatim", "constant"]


@pytest.mark.parametrize("bps,ch", [(8, 1), (12, 2), (16, 1), (16, 2), (20, 2), (24, 1), (24, 2)])
def test_roundtrip_all_subframe_kinds(tmp_path, bps, ch):
    rng = np.random.default_rng(bps * 10 + ch)
    blocksize = 1152
    n = blocksize * 40 + 333  # a short last block
    pcm = _signal(rng, n, ch, bps)
    pcm[blocksize * 7:blocksize * 8] = pcm[blocksize * 7]       # a constant block
    pcm[blocksize * 9:blocksize * 10] &= ~7                       # a block with three wasted bits
    orders = list(range(1, 33))

    def plan(fi):
        return dict(kind="constant" if fi == 7 else KINDS[fi % 7], stereo=[None, 8, 9, 10][fi % 4] if ch == 2 else None,
                    lpc_order=orders[(fi * 5) % 32], method=fi % 2, porder=[0, 1, 3, 5][fi % 4 if blocksize % 32 == 0 else 0],
                    escape_parts=(0,) if fi % 6 == 5 else (), wasted=3 if fi == 9 else 0)

    path = tmp_path / "t.flac"
    path.write_bytes(encode(pcm, bps, 22050, blocksize, plan, seed=bps + ch))
    got, n_frames, channels, rate, bits = read_pcm_file(path)
    assert (n_frames, channels, rate, bits) == (n, ch, 22050, bps)
    assert np.array_equal(got.reshape(-1, ch), pcm)


def test_every_lpc_order_narrow_and_wide(tmp_path):
    for bps in (16, 24):
        rng = np.random.default_rng(bps)
        blocksize = 512
        pcm = _signal(rng, blocksize * 32, 2, bps)
        path = tmp_path / f"o{bps}.flac"
        path.write_bytes(encode(pcm, bps, 44100, blocksize, lambda fi: dict(kind="lpc", lpc_order=fi + 1, stereo=10, porder=2), seed=bps))
        got = read_pcm_file(path)[0]
        assert np.array_equal(got.reshape(-1, 2), pcm), bps


def test_threaded_decode_equals_sequential(tmp_path, monkeypatch):
    """Frames are independent: a large stream is decoded on several threads (flac_reader.c: decode_flac_parallel, chain of
    frames found by header CRC and coded numbers). Same samples as the sequential decoder, for intact files and - through
    the fall-back - for a file with a damaged frame and one with foreign bytes spliced in."""
    import os
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    files = [os.path.join(golden, "song.flac"), os.path.join(golden, "song_s32.flac")]
    blob = bytearray(open(files[0], "rb").read())
    blob[len(blob) // 2] ^= 0x55                       # one damaged frame: dropped by both decoders
    (tmp_path / "damaged.flac").write_bytes(bytes(blob))
    blob = bytearray(open(files[0], "rb").read())
    blob[200000:200000] = bytes(range(256)) * 8         # 2 KB of foreign data between / inside frames
    (tmp_path / "spliced.flac").write_bytes(bytes(blob))
    files += [str(tmp_path / "damaged.flac"), str(tmp_path / "spliced.flac")]
    for path in files:
        monkeypatch.setenv("BLX_DECODE_THREADS", "1")
        ref = read_pcm_file(path)
        for threads in ("2", "8"):
            monkeypatch.setenv("BLX_DECODE_THREADS", threads)
            got = read_pcm_file(path)
            assert got[1:] == ref[1:] and np.array_equal(got[0], ref[0]), (path, threads)


def test_device_decoder_logic_on_the_host(tmp_path, monkeypatch):
    """The device FLAC decoder (csrc/flacdec.cu: one thread per frame, planar scratch, plain LPC loop, byte-wise CRC) is
    compiled for the host as well; BLX_FLAC_EMULATE routes the reader's accelerator hook to that instance. Same samples as
    the host decoder on the reference's fixtures and on encoder streams of every kind; a damaged file is refused by it and
    decoded by the fall-back."""
    import os
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    files = [os.path.join(golden, n) for n in ("song.flac", "song_s32.flac", "song_s32_mono.flac")]
    rng = np.random.default_rng(99)
    for bps, ch in ((16, 2), (24, 2), (12, 1)):
        blocksize = 1152
        pcm = _signal(rng, blocksize * 200, ch, bps)  # > 256 KB of frames: the chain scan is used

        def plan(fi):
            return dict(kind=KINDS[fi % 7], stereo=[None, 8, 9, 10][fi % 4] if ch == 2 else None, lpc_order=1 + (fi * 5) % 32,
                        method=fi % 2, porder=[0, 1, 3, 5][fi % 4], escape_parts=(0,) if fi % 6 == 5 else ())

        path = tmp_path / f"e{bps}_{ch}.flac"
        path.write_bytes(encode(pcm, bps, 44100, blocksize, plan, seed=bps))
        files.append(str(path))
    blob = bytearray(open(files[0], "rb").read())
    blob[len(blob) // 2] ^= 0x55
    (tmp_path / "damaged.flac").write_bytes(bytes(blob))
    files.append(str(tmp_path / "damaged.flac"))
    import ctypes
    import bliss_b200
    count = ctypes.CDLL(bliss_b200.LIB_PATH).blx_flac_accelerated_count
    for path in files:
        monkeypatch.setenv("BLX_FLAC_GPU", "0")
        ref = read_pcm_file(path)
        monkeypatch.setenv("BLX_FLAC_GPU", "1")
        monkeypatch.setenv("BLX_FLAC_GPU_MIN_SAMPLES", "0")
        monkeypatch.setenv("BLX_FLAC_EMULATE", "1")
        before = count()
        got = read_pcm_file(path)
        assert count() == before + (0 if "damaged" in path else 1), path  # the accelerator ran (and refused the damaged file)
        assert got[1:] == ref[1:] and np.array_equal(got[0], ref[0]), path


def test_repeated_stream_helper(tmp_path, monkeypatch):
    """flac_util.repeat_flac (renumbered, re-checksummed copies of a file's frames; what the GPU tests and probes use for
    minutes of audio) decodes to the original PCM repeated, on the threaded host decoder and on the emulated device one."""
    import os
    from flac_util import repeat_flac
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "song.flac")
    monkeypatch.setenv("BLX_FLAC_GPU", "0")
    one = read_pcm_file(src)
    (tmp_path / "r3.flac").write_bytes(repeat_flac(src, 3))
    got = read_pcm_file(tmp_path / "r3.flac")
    assert got[1] == 3 * one[1] and np.array_equal(got[0], np.tile(one[0], 3))
    monkeypatch.setenv("BLX_FLAC_GPU", "1")
    monkeypatch.setenv("BLX_FLAC_GPU_MIN_SAMPLES", "0")
    monkeypatch.setenv("BLX_FLAC_EMULATE", "1")
    assert np.array_equal(read_pcm_file(tmp_path / "r3.flac")[0], got[0])
