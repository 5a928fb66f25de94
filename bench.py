#!/usr/bin/env python
"""bench.py — songs/sec of the bl_analyze() hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our engine (CUDA sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref)

Workload (config.workload): the per-GPU share of BASELINE.json configs[2] — synthetic 3-minute
44.1 kHz mono float32 songs through the FULL bl_analyze pipeline (front-end, amplitude, frequency,
envelope/tempo/attack, rating). One "step" = one pass of the pipeline over one batch of
`--songs-per-step` songs that is already resident in HBM (65 GB at the default 2048, far larger
than the 126 MB L2, so no flush is needed between steps). 4 steps of 2048 = the 8 192 songs one GPU
owns in configs[2]. Songs shard across ranks with no data-path collective ("scaling": "weak").

One JSON line on stdout (rank 0). Beyond the base contract it carries
  roofline          the dominant kernel of the step (by device time, CUDA events on the launch stream)
  roofline_kernels  every kernel of the step: share, achieved GB/s of its algorithmic bytes and
                    fraction of the measured HBM peak; FP64 TFLOP/s for the envelope kernel
  spectral_only     BASELINE.json configs[1]: 1 024 x 30-s songs, fused front-end + Hann + rFFT-512 +
                    band-ratio kernel alone, with its HBM-read roofline fraction
  e2e               the same metric through the host-buffer C-ABI call (blx_analyze_batch_f32):
                    pinned host PCM -> device copies and the device -> host read of the results are
                    inside the timed region
  cpu_baseline      the reference's analyser sources (oracle/_ref, compiled verbatim with our shim FFT
                    in place of fftw3 / av_rdft) on a bounded sample of the same songs, all host cores
  parity            the GPU results of that sample against the CPU results (1e-4 relative)

oracle/ is used here only by the cpu_baseline / --impl reference legs (as the thing that is the
CPU baseline) — never on the measured GPU path.
"""
import argparse
import ctypes
import json
import math
import mmap
import os
import random
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RATE_IN = 44100
METRIC = "songs/sec (3-min 44.1kHz f32), full bl_analyze pipeline"
UNIT = "songs/s"
# FP64 work of the envelope kernel per hop: continuous FIR 25 flop x 256 samples,
# 512-point real FFT (256-point complex + split) ~ 13.3 kflop, power + float accumulation ~ 1 kflop
FP64_FLOP_PER_HOP = 25 * 256 + 13300 + 1028
DFMA_PEAK_TFLOPS = 36.4  # measured on this pool's B200 with tools/ubench.cu (profiles/r1_ubench.txt)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------ synthetic songs
def song_params(index):
    """Host-side, device-independent parameters of synthetic song `index`."""
    r = random.Random(0x5EED0000 + index)
    return dict(
        seed=0x5EED0000 + index,
        sigma=r.uniform(0.02, 0.12),
        tones=[(r.uniform(60.0, 8000.0), r.uniform(0.01, 0.08), r.uniform(0.0, 6.28)) for _ in range(r.randint(2, 4))],
        bpm=r.uniform(60.0, 180.0), decay=r.uniform(4.0, 12.0), burst_amp=r.uniform(0.05, 0.25),
        burst_f=r.uniform(80.0, 400.0),
    )


def synth_song(torch, out, index, t):
    """Fills `out` (float32 tensor, n samples, any device) with song `index`: band-limited noise +
    2-4 sinusoids + amplitude-modulated bursts at a per-song tempo; zero-mean, peak <= 0.5."""
    p = song_params(index)
    n = out.numel()
    g = torch.Generator(device=out.device)
    g.manual_seed(p["seed"])
    noise = torch.randn(n + 2, generator=g, device=out.device, dtype=torch.float32)
    sig = (noise[:-2] + 2.0 * noise[1:-1] + noise[2:]) * (p["sigma"] / math.sqrt(6.0))
    for f, a, ph in p["tones"]:
        sig += (a * torch.sin(2.0 * math.pi * f * t + ph)).to(torch.float32)
    phase = torch.remainder(t * (p["bpm"] / 60.0), 1.0)
    sig += (p["burst_amp"] * torch.exp(-phase * p["decay"]) * torch.sin(2.0 * math.pi * p["burst_f"] * t)).to(torch.float32)
    sig -= sig.mean()
    peak = float(sig.abs().max())
    if peak > 0.5:
        sig *= 0.5 / peak
    out.copy_(sig)


# ------------------------------------------------------------------------------------------ CPU leg
_CPU = {}


def _cpu_frontend(i):
    from oracle.binding import Oracle
    orc = _CPU.setdefault("orc", Oracle())
    _CPU["s16"][i, :] = orc.frontend_f32(_CPU["f32"][i])
    return i


def _cpu_analyze(i):
    kind = _CPU["kind"]
    dur = _CPU["duration"]
    if kind == "reference":
        from oracle.binding import RefLib
        ref = _CPU.setdefault("ref", RefLib())
        r = ref.analyze_pcm(_CPU["s16"][i], dur)
    else:
        from oracle.binding import Oracle
        orc = _CPU.setdefault("orc", Oracle())
        r = orc.analyze(_CPU["s16"][i], dur)
    return (i, r["tempo"], r["amplitude"], r["frequency"], r["attack"])


def _cpu_energy(i):
    from oracle.binding import Oracle
    orc = _CPU.setdefault("orc", Oracle())
    return (i, orc.envelope_energy(_CPU["s16"][i]))


def cpu_leg(pcm, steps, warmup, workers=None, energies=0, energy_path=None):
    """Times the reference's CPU analysers on the songs of `pcm`: S x n float32 (44.1 kHz mono) or S x n int16
    (22 050 Hz stereo interleaved, the analysers' native input).
    For float32 the f32 -> int16/22 050 Hz/stereo front-end (oracle/frontend.c, the same arithmetic the GPU fuses
    into pass 1) runs first and is NOT timed: decode/resample is excluded on both sides.
    Worker PROCESSES, one per core: the reference is not thread-safe (fftw planner).
    energies = K > 0: also writes the hop energies E[m] (oracle restatement) of the first K songs to energy_path."""
    import multiprocessing as mp

    from oracle.binding import REF_SO
    S, n = pcm.shape
    cores = os.cpu_count() or 1
    workers = max(1, min(workers or cores, S))
    is_f32 = pcm.dtype == np.float32
    n16 = 2 * (n // 2) if is_f32 else n
    if is_f32:
        shm = mmap.mmap(-1, S * n16 * 2)
        _CPU["f32"] = pcm
        _CPU["s16"] = np.frombuffer(shm, dtype=np.int16).reshape(S, n16)
        _CPU["duration"] = n // RATE_IN
    else:
        _CPU["s16"] = pcm
        _CPU["duration"] = n // RATE_IN  # n int16 values = n / 2 frames at 22 050 Hz
    _CPU["kind"] = "reference" if os.path.exists(REF_SO) else "port"
    ctx = mp.get_context("fork")
    times, results = [], None
    with ctx.Pool(workers) as pool:
        if is_f32:
            pool.map(_cpu_frontend, range(S), chunksize=1)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            results = pool.map(_cpu_analyze, range(S), chunksize=1)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
        if energies > 0 and energy_path:
            en = pool.map(_cpu_energy, range(min(energies, S)), chunksize=1)
            np.save(energy_path, np.stack([e for _, e in sorted(en, key=lambda t: t[0])]))
    total = sum(times)
    res = np.zeros((S, 4), dtype=np.float64)
    for i, a, b, c, d in results:
        res[i] = (a, b, c, d)
    return dict(value=S * len(times) / total, seconds_per_step=total / len(times), workers=workers, cores=cores,
                kind=_CPU["kind"], songs=S, results=res.tolist())


def cpu_leg_subprocess(pcm, steps, warmup, energies=0):
    """Runs cpu_leg in a fresh interpreter (no CUDA context to fork). Returns (leg dict, energies or None)."""
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        path = os.path.join(d, "sample.npy")
        epath = os.path.join(d, "energy.npy")
        np.save(path, pcm)
        cmd = [sys.executable, os.path.abspath(__file__), "--_cpu-leg", path, "--steps", str(steps), "--warmup", str(warmup)]
        if energies:
            cmd += ["--_cpu-energies", str(energies), "--_cpu-energy-path", epath]
        out = subprocess.run(cmd, check=True, capture_output=True, text=True)
        E = np.load(epath) if energies else None
    return json.loads(out.stdout.strip().splitlines()[-1]), E


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons, power = [], 0.0, set(), 0.0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1]))
                smax = max(smax, float(c[2]))
                power = max(power, float(c[3]))
            except ValueError:
                continue
            for nm, v in zip(names, c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return None
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=smax, reasons=sorted(reasons), power_w_max=power,
                    samples=len(sm))


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    n_in = int(args.seconds * RATE_IN)
    cores = os.cpu_count() or 1
    S = max(2, min(4 * cores, 128))
    t = torch.arange(n_in, dtype=torch.float64) / RATE_IN
    f32 = np.zeros((S, n_in), dtype=np.float32)
    for i in range(S):
        synth_song(torch, torch.from_numpy(f32[i]), i, t)
    r = cpu_leg(f32, args.steps, args.warmup)
    sample = (f"{S} synthetic {args.seconds:g}-s songs per step ({r['workers']} worker processes, decode/front-end "
              f"excluded); analysers = reference src/*.c compiled verbatim, shim FFT in place of fftw3/av_rdft"
              if r["kind"] == "reference" else f"{S} songs per step, oracle C port")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32 (CPU)", "data": "synthetic",
        # the SAME config as our arm (the workload being compared); each step here is a bounded sample of it (cpu_baseline.sample)
        "config": workload_config(args, args.songs_per_step, args.gpus),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["workers"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, songs_per_step, n_gpus):
    return {
        "workload": f"BASELINE.json configs[2], per-GPU share: synthetic {args.seconds:g}-s 44.1 kHz mono float32 songs, "
                    f"full bl_analyze pipeline (front-end + amplitude + frequency + envelope/tempo/attack + rating)",
        "songs_per_step_per_gpu": songs_per_step, "song_seconds": args.seconds, "samples_per_song": int(args.seconds * RATE_IN),
        "bytes_per_song": int(args.seconds * RATE_IN) * 4, "sharding": f"songs x {n_gpus} ranks, no data-path collective",
        "l2": "inputs (GBs per step) far exceed the 126 MB L2; no flush needed",
    }


# ------------------------------------------------------------------------------------------ parity campaign
def make_s16_stereo(torch, left_f32, right_f32):
    """int16 / 22 050 Hz / stereo interleaved from two 44.1 kHz float32 songs: L = every other sample of the
    first, R = a mix of both (decorrelated channels, the reference decoder's usual output)."""
    L = torch.clamp(torch.round(left_f32[::2] * 32768.0), -32768, 32767)
    R = torch.clamp(torch.round((0.6 * left_f32[::2] + 0.4 * right_f32[::2]) * 32768.0), -32768, 32767)
    return torch.stack([L, R], dim=1).reshape(-1).to(torch.int16)


def parity_campaign(torch, eng, E, dev, stream, rank, world, total, dist):
    """GPU records against the CPU analysers (oracle/_ref when built, else the oracle port) over `total` songs
    shared by all ranks: 30-s and 3-minute songs, float32 (front-end + doubled-mono kernels) AND native int16
    stereo with decorrelated channels (the reference's real input). Reports onset-count mismatches (must be 0),
    the largest relative error per component and the measured rate of single-ulp flips of E[m]."""
    import bliss_b200
    per_rank = -(-total // world)
    n_long = max(2, per_rank // 32)  # 3-minute songs of each format
    n_short = max(2, (per_rank - 2 * n_long + 1) // 2)
    groups = [("f32_30s", "f32", 30, n_short, 256), ("s16_30s", "s16", 30, n_short, 256),
              ("f32_180s", "f32", 180, n_long, 32), ("s16_180s", "s16", 180, n_long, 32)]
    comps = ("tempo", "amplitude", "frequency", "attack")
    max_rel = {k: 0.0 for k in comps}
    beat_mismatch, status_bad, n_done, cpu_s = 0, 0, 0, 0.0
    flips, hops, e_max_rel = 0, 0, 0.0
    per_group = {}
    kind = None
    base_index = 1_000_000 + rank * (per_rank + 8)
    song_i = 0
    for name, fmt, seconds, count, chunk in groups:
        n_in = seconds * RATE_IN
        t = torch.arange(n_in, dtype=torch.float64, device=dev) / RATE_IN
        g_rel, g_beat = 0.0, 0
        first_chunk = True
        for c0 in range(0, count, chunk):
            S = min(chunk, count - c0)
            f32 = torch.empty((S + 1, n_in), dtype=torch.float32, device=dev)
            for i in range(S + 1):
                synth_song(torch, f32[i], base_index + song_i + i, t)
            song_i += S
            if fmt == "f32":
                stride = (n_in + 63) // 64 * 64 + 64
                buf = torch.zeros(S * stride, dtype=torch.float32, device=dev)
                buf.view(S, stride)[:, :n_in] = f32[:S]
                host = f32[:S].cpu().numpy()
                d_out = torch.zeros(S * 8, dtype=torch.int32, device=dev)
                eng.analyze_device(E.FMT_F32, buf.data_ptr(), [i * stride for i in range(S)], [n_in] * S, d_out.data_ptr(),
                                   stream=stream)
            else:
                n16 = 2 * (n_in // 2)
                stride = (n16 + 63) // 64 * 64 + 64
                buf = torch.zeros(S * stride, dtype=torch.int16, device=dev)
                for i in range(S):
                    buf[i * stride:i * stride + n16] = make_s16_stereo(torch, f32[i], f32[i + 1])
                host = buf.view(S, stride)[:, :n16].cpu().numpy()
                d_out = torch.zeros(S * 8, dtype=torch.int32, device=dev)
                eng.analyze_device(E.FMT_S16, buf.data_ptr(), [i * stride for i in range(S)], [n16] * S, d_out.data_ptr(),
                                   durations=[seconds] * S, stream=stream)
            torch.cuda.synchronize()
            got = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=bliss_b200.RESULT_DTYPE)
            del buf, f32
            n_e = 16 if (first_chunk and seconds == 30) else (2 if first_chunk else 0)
            r, Eref = cpu_leg_subprocess(host, 1, 0, energies=n_e)
            kind = r["kind"]
            cpu_s += r["seconds_per_step"]
            ref = np.array(r["results"])
            status_bad += int(np.count_nonzero(got["status"]))
            for j, k in enumerate(comps):
                rel = np.abs(got[k].astype(np.float64) - ref[:, j]) / np.maximum(np.abs(ref[:, j]), 1e-30)
                max_rel[k] = max(max_rel[k], float(rel.max()))
                g_rel = max(g_rel, float(rel.max()))
            # tempo = 4 * beat / duration - 30.4 in float: equal tempo <=> equal onset count
            ref_beat = np.rint((ref[:, 0] + 30.4) * seconds / 4.0).astype(np.int64)
            bad = int(np.count_nonzero((got["beat"] != ref_beat) | (got["tempo"] != ref[:, 0].astype(np.float32))))
            beat_mismatch += bad
            g_beat += bad
            for i in range(n_e):  # hop energies of a few songs: how often does the float accumulator land one ulp off?
                Eg = eng.envelope_energy_f32(host[i]) if fmt == "f32" else eng.envelope_energy(host[i])
                flips += int(np.count_nonzero(Eg != Eref[i]))
                hops += int(Eg.size)
                e_max_rel = max(e_max_rel, float(np.max(np.abs(Eg - Eref[i]) / np.maximum(Eref[i], 1e-300))))
            n_done += S
            first_chunk = False
        per_group[name] = {"songs": count, "max_rel_err": g_rel, "beat_mismatches": g_beat}
        del t
    tot = torch.tensor([n_done, beat_mismatch, status_bad, flips, hops], dtype=torch.float64, device=dev)
    mx = torch.tensor([max_rel[k] for k in comps] + [e_max_rel, cpu_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    tot, mx = tot.tolist(), mx.tolist()
    worst = max(mx[:4])
    return {"songs": int(tot[0]), "ranks": world, "checker": kind, "tolerance": 1e-4,
            "beat_mismatches": int(tot[1]), "status_nonzero": int(tot[2]),
            "max_rel_err": {k: mx[j] for j, k in enumerate(comps)},
            "ok": bool(worst <= 1e-4 and tot[1] == 0 and tot[2] == 0),
            "energy_hops_compared": int(tot[4]), "energy_flips": int(tot[3]),
            "energy_flip_rate": (tot[3] / tot[4]) if tot[4] else None, "energy_max_rel_err": mx[4],
            "groups_rank0": per_group, "cpu_seconds_max_rank": mx[5],
            "mix": "per rank: 30-s and 3-min songs, half 44.1 kHz float32 (front-end path), half native int16 stereo with a "
                   "decorrelated right channel; tempo compared exactly (<=> onset count), E[m] bit for bit on a sample"}


# ------------------------------------------------------------------------------------------ configs[4]
def chain_leg(torch, eng, E, dev, stream, rank, world, dist, barrier, max_over_ranks, buf, stride, n_in, B, per_gpu,
              hbm_peak):
    """BASELINE.json configs[4], chained: every rank analyses its share of world x per_gpu distinct synthetic songs in
    batches of B (generated on the device, untimed, into the resident buffer), the force vectors are all-gathered
    (NCCL, 16 B/song), and every rank computes its row slab of the all-pairs bl_distance matrix twice: fused nearest-
    neighbour epilogue, and materialised in HBM (the write-bound form). Times are CUDA events, max over ranks."""
    import bliss_b200
    from bliss_b200 import parallel
    n_batches = max(1, per_gpu // B)
    per_gpu = n_batches * B
    t = torch.arange(n_in, dtype=torch.float64, device=dev) / RATE_IN
    d_res = torch.zeros(per_gpu * 8, dtype=torch.int32, device=dev)
    offs, lens = [i * stride for i in range(B)], [n_in] * B
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    first = 2_000_000 + rank * per_gpu
    analysis_ms, gen_s = 0.0, 0.0
    for b in range(n_batches):
        t0 = time.perf_counter()
        for i in range(B):
            synth_song(torch, buf[i * stride:i * stride + n_in], first + b * B + i, t)
        torch.cuda.synchronize()
        gen_s += time.perf_counter() - t0
        ev0.record()
        eng.analyze_device(E.FMT_F32, buf.data_ptr(), offs, lens, d_res.data_ptr() + b * B * 32, stream=stream)
        ev1.record()
        torch.cuda.synchronize()
        analysis_ms += ev0.elapsed_time(ev1)
    analysis_ms = max_over_ranks(analysis_ms)
    rec = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=bliss_b200.RESULT_DTYPE)
    bad = int(np.count_nonzero(rec["status"]))
    local = d_res.view(torch.float32).view(per_gpu, 8)[:, :4].contiguous()
    parallel.all_gather_vectors(local)  # NCCL channel set-up for this message size is not part of the measurement
    barrier()
    ev0.record()
    allv, row0 = parallel.all_gather_vectors(local)
    ev1.record()
    barrier()
    gather_ms = max_over_ranks(ev0.elapsed_time(ev1))
    n_total = allv.shape[0]
    # the gathered table against an independent re-analysis: this rank re-generates the first songs of the NEXT rank's
    # shard and analyses them itself (at world == 1: its own first songs, alone in a small batch)
    K = min(64, B)
    peer = (rank + 1) % world
    for i in range(K):
        synth_song(torch, buf[i * stride:i * stride + n_in], 2_000_000 + peer * per_gpu + i, t)
    d_chk = torch.zeros(K * 8, dtype=torch.int32, device=dev)
    eng.analyze_device(E.FMT_F32, buf.data_ptr(), offs[:K], lens[:K], d_chk.data_ptr(), stream=stream)
    torch.cuda.synchronize()
    mine = d_chk.view(torch.float32).view(K, 8)[:, :4]
    same = bool(torch.equal(mine.view(torch.int32), allv[peer * per_gpu:peer * per_gpu + K].view(torch.int32)))
    # fused nearest-neighbour epilogue over this rank's rows
    idx = torch.empty(per_gpu, dtype=torch.int32, device=dev)
    dst = torch.empty(per_gpu, dtype=torch.float32, device=dev)
    eng.distance_nearest_device(allv.data_ptr(), n_total, row0, per_gpu, idx.data_ptr(), dst.data_ptr(), 0, stream=stream)
    barrier()
    ev0.record()
    eng.distance_nearest_device(allv.data_ptr(), n_total, row0, per_gpu, idx.data_ptr(), dst.data_ptr(), 0, stream=stream)
    ev1.record()
    barrier()
    near_ms = max_over_ranks(ev0.elapsed_time(ev1))
    # materialised slab (per_gpu x n_total float32 in HBM)
    slab = torch.empty((per_gpu, n_total), dtype=torch.float32, device=dev)
    eng.distance_rows_device(allv.data_ptr(), n_total, row0, per_gpu, slab.data_ptr(), stream=stream)
    barrier()
    ev0.record()
    eng.distance_rows_device(allv.data_ptr(), n_total, row0, per_gpu, slab.data_ptr(), stream=stream)
    ev1.record()
    barrier()
    slab_ms = max_over_ranks(ev0.elapsed_time(ev1))
    slab_bytes = per_gpu * n_total * 4
    # the fused epilogue against a brute-force scan of the materialised rows: same distance bit for bit, and no
    # lower index at that distance (ties go to the lowest index, as a scan over bl_distance values would)
    ok_near = True
    rows_per = max(1, (1 << 28) // max(n_total, 1))
    cols = torch.arange(n_total, device=dev, dtype=torch.int64)
    for r0 in range(0, per_gpu, rows_per):
        r1 = min(per_gpu, r0 + rows_per)
        blk = slab[r0:r1].clone()
        blk[torch.arange(r1 - r0, device=dev), torch.arange(row0 + r0, row0 + r1, device=dev)] = float("inf")
        mn = blk.min(dim=1).values
        first_idx = torch.where(blk == mn[:, None], cols[None, :], n_total).min(dim=1).values
        ok_near = ok_near and bool(torch.equal(mn, dst[r0:r1])) and bool(torch.equal(first_idx, idx[r0:r1].long()))
        del blk
    flags = torch.tensor([1.0 if same else 0.0, 1.0 if ok_near else 0.0, float(-bad)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    flags = flags.tolist()
    del slab, allv
    songs_total = world * per_gpu
    total_ms = analysis_ms + gather_ms + near_ms
    return {"workload": f"BASELINE.json configs[4]: {songs_total} distinct synthetic 3-min songs end to end on {world} GPU(s) "
                        f"({per_gpu} per GPU in batches of {B}) -> all-gather of the force vectors -> all-pairs bl_distance",
            "songs": songs_total, "analysis_ms": analysis_ms, "analysis_songs_per_s": songs_total / (analysis_ms * 1e-3),
            "all_gather_ms": gather_ms, "all_gather_bytes": songs_total * 16,
            "nearest_ms": near_ms, "nearest_pairs_per_s": float(per_gpu) * n_total * world / (near_ms * 1e-3),
            "chained_ms": total_ms, "chained_songs_per_s": songs_total / (total_ms * 1e-3),
            "slab_ms": slab_ms, "slab_bytes_per_gpu": slab_bytes, "slab_gbs": slab_bytes / (slab_ms * 1e-3) / 1e9,
            "slab_frac_hbm_write": slab_bytes / (slab_ms * 1e-3) / 1e9 / hbm_peak,
            "gathered_vectors_bit_equal_reanalysis": bool(flags[0] == 1.0), "reanalysed_songs_per_rank": K,
            "nearest_equals_bruteforce_over_slab": bool(flags[1] == 1.0), "status_nonzero": int(-flags[2]),
            "generation_s_untimed": gen_s}


# ------------------------------------------------------------------------------------------ drop-in path
def bl_analyze_leg(torch, buf, stride, n_in, n_files=8):
    """songs/s through the reference's own entry point, bl_analyze(filename, &song) of include/bliss.h: file read + decode
    on the host, analysis on the GPU, per call; from 1 caller thread and from 8 (each call takes an engine out of the
    library's pool, so concurrent callers overlap on the GPU). Files: the reference's 11-s fixture, 3-minute WAVs in
    the analysers' native format (int16 / 22 050 Hz / stereo, written from the step's first songs) and 3-minute CD-audio
    WAVs (44.1 kHz / 16 bit / stereo: through the decode-stage resampler)."""
    import struct
    import threading

    import bliss_b200
    L = bliss_b200.load()
    out = {}
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        wavs = []
        n16 = 2 * (n_in // 2)
        for i in range(n_files):
            mono = torch.clamp(torch.round(buf[i * stride:i * stride + n_in:2][:n16 // 2] * 32768.0), -32768, 32767).to(torch.int16)
            pcm = torch.stack([mono, torch.roll(mono, 3)], dim=1).reshape(-1).cpu().numpy()
            raw = pcm.tobytes()
            path = os.path.join(d, f"song{i}.wav")
            with open(path, "wb") as f:
                f.write(b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " +
                        struct.pack("<IHHIIHH", 16, 1, 2, 22050, 22050 * 4, 4, 16) + b"data" + struct.pack("<I", len(raw)) + raw)
            wavs.append(path.encode())
        # the most common real-world input: CD audio (44.1 kHz / 16 bit / stereo) -> decode-stage resampler -> analysis
        cd_wavs = []
        for i in range(4):
            mono = torch.clamp(torch.round(buf[i * stride:i * stride + n_in] * 32768.0), -32768, 32767).to(torch.int16)
            pcm = torch.stack([mono, torch.roll(mono, 5)], dim=1).reshape(-1).cpu().numpy()
            raw = pcm.tobytes()
            path = os.path.join(d, f"cd{i}.wav")
            with open(path, "wb") as f:
                f.write(b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " +
                        struct.pack("<IHHIIHH", 16, 1, 2, 44100, 44100 * 4, 4, 16) + b"data" + struct.pack("<I", len(raw)) + raw)
            cd_wavs.append(path.encode())
        fixture = os.path.join(ROOT, "tests", "golden", "song.flac").encode()
        # three minutes of FLAC: the fixture's frames 16 times over (22 050 Hz: device decode, native format) and six
        # seconds of CD audio from the test encoder 30 times over (44.1 kHz: device decode + resampler fused)
        flacs = {}
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from flac_encode import encode
            from flac_util import repeat_flac
            with open(os.path.join(d, "long.flac"), "wb") as f:
                f.write(repeat_flac(fixture.decode(), 16))
            flacs["flac_3min_22k"] = os.path.join(d, "long.flac").encode()
            six = buf[:6 * 44100].cpu().numpy().astype(np.float64)
            xs = np.round(np.stack([six * 0.9, np.roll(six, 23) * 0.6], axis=1) * 32767).astype(np.int64)
            with open(os.path.join(d, "six.flac"), "wb") as f:
                f.write(encode(xs, 16, 44100, 4096, lambda fi: dict(kind="lpc", stereo=10, lpc_order=8, porder=3), seed=3))
            with open(os.path.join(d, "cd.flac"), "wb") as f:
                f.write(repeat_flac(os.path.join(d, "six.flac"), 30))
            flacs["flac_3min_cd_44k"] = os.path.join(d, "cd.flac").encode()
        except Exception as ex:  # the leg is informative: never fail the bench over a helper
            flacs = {"error": repr(ex)}

        def run(files, threads, reps):
            recs, lock = {}, threading.Lock()

            def work(tid):
                for r in range(reps):
                    for k in range(tid, len(files), threads):
                        s = bliss_b200.BlSong()
                        rc = L.bl_analyze(files[k], ctypes.byref(s))
                        rec = (rc, s.force, s.force_vector.tempo, s.force_vector.amplitude, s.force_vector.frequency, s.force_vector.attack)
                        L.bl_free_song(ctypes.byref(s))
                        with lock:
                            recs.setdefault(k, set()).add(rec)
            ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
            t0 = time.perf_counter()
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            dt = time.perf_counter() - t0
            return len(files) * reps / dt, recs

        def median_of(files, threads, reps, trials=3):
            # calls of a few milliseconds from several threads are at the mercy of the host's allocator and scheduler:
            # the median of three short trials is reported, results of all of them are kept for the identity check
            rates, recs = [], {}
            for _ in range(trials):
                r, rc = run(files, threads, reps)
                rates.append(r)
                for k, v in rc.items():
                    recs.setdefault(k, set()).update(v)
            return sorted(rates)[len(rates) // 2], recs

        run(wavs[:2], 1, 1)  # warm-up: engine pool, kernel attributes
        run(wavs, 8, 1)
        one, r1 = median_of(wavs, 1, 1)
        many, r8 = median_of(wavs, 8, 4)
        same = all(len(v) == 1 for v in r1.values()) and all(len(v) == 1 for v in r8.values()) and all(r1[k] == r8[k] for k in r1)
        fx1, _ = median_of([fixture] * 8, 1, 2)
        fx8, _ = median_of([fixture] * 8, 8, 4)
        run(cd_wavs[:1], 1, 1)
        cd1, _ = median_of(cd_wavs, 1, 1)
        cd4, _ = median_of(cd_wavs, 4, 2)
        flac_rates = {}
        for name, path in flacs.items():
            if name == "error":
                flac_rates[name] = path
                continue
            run([path], 1, 1)
            flac_rates[name + "_songs_per_s_1_thread"], _ = median_of([path] * 4, 1, 1)
            flac_rates[name + "_songs_per_s_4_threads"], _ = median_of([path] * 4, 4, 2)
            os.environ["BLX_FLAC_GPU"] = "0"  # the same with the four host decode threads instead of the device decoder
            flac_rates[name + "_host_decode_songs_per_s_1_thread"], _ = median_of([path] * 4, 1, 1)
            del os.environ["BLX_FLAC_GPU"]
        out = {"api": "bl_analyze (include/bliss.h), one file per call: host read + decode, GPU analysis",
               "wav_3min_songs_per_s_1_thread": one, "wav_3min_songs_per_s_8_threads": many,
               "fixture_11s_songs_per_s_1_thread": fx1, "fixture_11s_songs_per_s_8_threads": fx8,
               "cd_wav_44k_3min_songs_per_s_1_thread": cd1, "cd_wav_44k_3min_songs_per_s_4_threads": cd4,
               "results_identical_across_threads": bool(same), "files": n_files, "statistic": "median of 3 trials"}
        out.update(flac_rates)
    return out


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch

    import bliss_b200
    from bliss_b200 import engine as E

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        tns = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns.item())

    eng = bliss_b200.Engine(local_rank)
    B = args.songs_per_step
    n_in = int(args.seconds * RATE_IN)
    stride = (n_in + 63) // 64 * 64 + 64
    buf = torch.zeros(B * stride, dtype=torch.float32, device=dev)
    t = torch.arange(n_in, dtype=torch.float64, device=dev) / RATE_IN
    t0 = time.perf_counter()
    for i in range(B):
        synth_song(torch, buf[i * stride:i * stride + n_in], rank * B + i, t)
    torch.cuda.synchronize()
    del t
    log(f"[rank {rank}] generated {B} songs ({B * n_in * 4 / 1e9:.1f} GB) in {time.perf_counter() - t0:.1f}s")
    offs = (ctypes.c_int64 * B)(*[i * stride for i in range(B)])  # built once: the same batch layout for every launch
    lens = (ctypes.c_int64 * B)(*([n_in] * B))
    d_out = torch.zeros(B * 8, dtype=torch.int32, device=dev)
    # All timed work and the CUDA events that bracket it go to ONE explicit (non-default) stream: the
    # C-ABI treats a NULL stream as "the engine's own", which torch events on the default stream
    # would not see.
    torch.cuda.synchronize()
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    # A step enqueues one batch; the job's batches are pipelined (the sequential tail of one batch runs under the next
    # batch's kernels) and joined once, inside the timed region.
    def step():
        eng.analyze_device(E.FMT_F32, buf.data_ptr(), offs, lens, d_out.data_ptr(), stream=stream, wait=False)

    # ---------------- device-resident steps: `value`
    eng.profile(True)
    for _ in range(args.warmup):
        step()
    eng.join(stream)
    barrier()
    fp64_peak = max_over_ranks(eng.measure_fp64_peak())  # DFMA roofline denominator, measured here and now
    barrier()
    eng.profile_reset()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    eng.join(stream)
    ev1.record()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if sampler else None
    launches = eng.launch_count() - launches0
    prof = eng.profile_read()
    eng.profile(False)
    res = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=bliss_b200.RESULT_DTYPE)
    bad = int(np.count_nonzero(res["status"]))
    if bad:
        raise SystemExit(f"bench.py: {bad} songs were not analysed (status != 0)")
    value = world * B * args.steps / (ms_total / 1e3)

    # ---------------- per-kernel rooflines (CUDA events around every launch, same stream)
    hbm_peak, peak_src = 6650.0, "fallback"
    try:
        hbm_peak, peak_src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        pass
    n_ms = n_in // 2
    hops = 2 * ((2 * n_ms) // 512) - 2
    alg_bytes = {  # per song, DESIGN.md §4
        "pass1_kernel": n_in * 4,              # one read of the float32 PCM
        "envelope_kernel": n_ms * 2,           # one read of the decimated int16 stream
        "epilogue_kernel": 3808 * 4 + 256 * 4 * 7,
        "tail_kernel": (hops + 2) * 8,
    }
    # The engine launches every kernel once per sub-batch of 512 songs (tail_kernel: log compression + tail, two
    # launches), so a "launch" below is the step's launches of that kernel taken together: B songs.
    kern_ms_sum = sum(v[0] for v in prof.values()) or 1.0
    kernels = {}
    for name, (ms, n) in prof.items():
        if n == 0:
            continue
        per_step_ms = ms / args.steps
        ach = alg_bytes.get(name, 0) * B / (per_step_ms * 1e-3) / 1e9
        kernels[name] = {"share": ms / kern_ms_sum, "ms_per_step": per_step_ms, "launches_per_step": n / args.steps,
                         "ms_per_launch": ms / n, "alg_bytes_per_step": alg_bytes.get(name, 0) * B, "achieved_gbs": ach,
                         "frac_hbm": ach / hbm_peak}
        if name == "envelope_kernel":
            tf = FP64_FLOP_PER_HOP * hops * B / (per_step_ms * 1e-3) / 1e12
            kernels[name]["fp64_tflops"] = tf
            kernels[name]["frac_fp64"] = tf / fp64_peak
    if "tail_kernel" in kernels:
        kernels["tail_kernel"]["note"] = ("runs on the engine's high-priority tail stream UNDER the next sub-batch's kernels: its "
                                          "time overlaps theirs (the shares add up to more than the step)")
    dom = max(kernels, key=lambda k: kernels[k]["share"])
    traffic = None
    try:
        per_song = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
        traffic = per_song * B if per_song else None  # bytes per launch (ncu capture scaled to this batch)
    except Exception:
        pass
    if dom == "envelope_kernel":
        roofline = {"kernel": dom, "bound": "fp64", "achieved": kernels[dom]["fp64_tflops"], "peak": fp64_peak,
                    "unit": "TFLOP/s", "frac": kernels[dom]["frac_fp64"], "traffic": traffic,
                    "peak_source": "DFMA throughput measured in this run before the timed steps (blx_measure_fp64_peak); "
                                   f"round-1 microbenchmark: {DFMA_PEAK_TFLOPS} (profiles/r1_ubench.txt)",
                    "alg_flop_per_step": FP64_FLOP_PER_HOP * hops * B, "ms_per_step": kernels[dom]["ms_per_step"],
                    "launches_per_step": kernels[dom]["launches_per_step"],
                    "note": "FP64-pipe bound (SURVEY.md §0 F5, §7.3 H3); its HBM fraction is in roofline_kernels"}
    else:
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": hbm_peak,
                    "unit": "GB/s", "frac": kernels[dom]["frac_hbm"], "traffic": traffic, "peak_source": peak_src,
                    "ms_per_step": kernels[dom]["ms_per_step"], "launches_per_step": kernels[dom]["launches_per_step"]}

    # the step's largest HBM-bound kernel in the contract's own form (bound "hbm", peak from MEASURED_PEAKS.json)
    roofline_hbm = None
    if "pass1_kernel" in kernels:
        k1 = kernels["pass1_kernel"]
        t1 = None
        try:
            per_song = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("pass1_kernel")
            t1 = per_song * B if per_song else None
        except Exception:
            pass
        roofline_hbm = {"kernel": "pass1_kernel<F32, FULL>", "bound": "hbm", "achieved": k1["achieved_gbs"], "peak": hbm_peak,
                        "unit": "GB/s", "frac": k1["frac_hbm"], "traffic": t1, "peak_source": peak_src,
                        "share_of_step": k1["share"],
                        "note": "algorithmic bytes = one read of the float32 PCM; the kernel also writes the 7.9 MB/song "
                                "decimated stream (traffic counts both)"}

    # ---------------- configs[1]: the fused spectral kernel alone, 1 024 x 30-s songs
    spectral = None
    if not args.no_spectral:
        n30 = 30 * RATE_IN
        Bs = min(B, 1024)
        if n30 <= n_in:
            d_freq = torch.zeros(Bs, dtype=torch.float32, device=dev)
            so, sl = (ctypes.c_int64 * Bs)(*offs[:Bs]), (ctypes.c_int64 * Bs)(*([n30] * Bs))
            def spectral_pass():
                eng.spectral_device(E.FMT_F32, buf.data_ptr(), so, sl, d_freq.data_ptr(), stream=stream)
            for _ in range(max(3, args.warmup)):
                spectral_pass()
            # (a) songs/s of the pass as a caller sees it: no per-kernel events in the stream
            eng.profile(False)
            barrier()
            reps = 20
            ev0.record()
            for _ in range(reps):
                spectral_pass()
            ev1.record()
            barrier()
            sp_ms = max_over_ranks(ev0.elapsed_time(ev1)) / reps
            # (b) the kernel's own duration for the roofline: CUDA events around every launch
            eng.profile(True)
            eng.profile_reset()
            for _ in range(10):
                spectral_pass()
            barrier()
            sp = eng.profile_read()
            eng.profile(False)
            k_ms = sp["pass1_kernel"][0] / max(sp["pass1_kernel"][1], 1)
            gbs = Bs * n30 * 4 / (k_ms * 1e-3) / 1e9
            spectral = {"workload": f"BASELINE.json configs[1]: {Bs} x 30-s 44.1 kHz mono f32 songs per GPU, fused front-end + "
                                    "Hann + rFFT-512 + per-bin power kernel + band-ratio epilogue only",
                        "value": world * Bs / (sp_ms * 1e-3), "unit": UNIT, "ms_per_pass": sp_ms,
                        "roofline": {"kernel": "pass1_kernel<F32, lite>", "bound": "hbm", "achieved": gbs, "peak": hbm_peak,
                                     "unit": "GB/s", "frac": gbs / hbm_peak, "ms_per_launch": k_ms,
                                     "alg_bytes_per_launch": Bs * n30 * 4, "peak_source": peak_src}}

    # ---------------- the analysers' native input (int16 / 22 050 Hz / stereo, what the reference's decoder hands over):
    # the same songs after the front-end, full pipeline without it; informational (the metric is the float32 workload)
    native = None
    if args.s16_songs > 0:
        Bn = min(args.s16_songs, B)
        n16 = 2 * (n_in // 2)                       # interleaved L,R
        stride16 = (n16 + 63) // 64 * 64 + 64
        s16 = torch.zeros(Bn * stride16, dtype=torch.int16, device=dev)
        for i in range(Bn):
            mono = torch.clamp(torch.round(buf[i * stride:i * stride + n_in:2][:n16 // 2] * 32768.0), -32768, 32767).to(torch.int16)
            view = s16[i * stride16:i * stride16 + n16].view(-1, 2)
            view[:, 0] = mono
            view[:, 1] = torch.roll(mono, 3)          # decorrelated right channel
        o16, l16 = (ctypes.c_int64 * Bn)(*[i * stride16 for i in range(Bn)]), (ctypes.c_int64 * Bn)(*([n16] * Bn))
        durs = (ctypes.c_uint64 * Bn)(*([int(args.seconds)] * Bn))
        d_out16 = torch.zeros(Bn * 8, dtype=torch.int32, device=dev)
        def step16():
            eng.analyze_device(E.FMT_S16, s16.data_ptr(), o16, l16, d_out16.data_ptr(), durations=durs, stream=stream, wait=False)
        for _ in range(3):
            step16()
        eng.join(stream)
        eng.profile(True)
        eng.profile_reset()
        barrier()
        ev0.record()
        for _ in range(3):
            step16()
        eng.join(stream)
        ev1.record()
        barrier()
        ms16 = max_over_ranks(ev0.elapsed_time(ev1)) / 3
        p16 = eng.profile_read()
        eng.profile(False)
        r16 = np.frombuffer(d_out16.cpu().numpy().tobytes(), dtype=bliss_b200.RESULT_DTYPE)
        native = {"workload": f"{Bn} x {args.seconds:g}-s int16 / 22 050 Hz / stereo songs per GPU (the reference decoder's output "
                              "format), full pipeline", "value": world * Bn / (ms16 * 1e-3), "unit": UNIT, "ms_per_pass": ms16,
                  "bytes_per_song": n16 * 2, "all_status_ok": bool(np.all(r16["status"] == 0)),
                  "kernel_ms_per_pass": {k: v[0] / 3 for k, v in p16.items() if v[1]},
                  "envelope_frac_fp64": (FP64_FLOP_PER_HOP * hops * Bn / (p16["envelope_kernel"][0] / 3 * 1e-3) / 1e12 / fp64_peak)
                  if p16.get("envelope_kernel", (0, 0))[1] else None}
        del s16

    # ---------------- configs[3]/[4]: all-pairs bl_distance over 1 M force vectors, fused nearest-neighbour
    # epilogue; vectors are sharded by rank, all-gathered (NCCL, 16 B/song), each rank does its row slab
    all_pairs = None
    if not args.no_distance:
        from bliss_b200 import parallel
        nv = args.distance_vectors
        gen = torch.Generator(device=dev)
        gen.manual_seed(0xD157)
        table = torch.randn((nv, 4), generator=gen, device=dev, dtype=torch.float32) * torch.tensor(
            [8.0, 6.0, 10.0, 12.0], device=dev)
        lo, hi = parallel.shard_range(nv, rank, world)
        local = table[lo:hi].contiguous()
        parallel.nearest_neighbours(eng, local[:min(hi - lo, 2048)])  # warm-up
        parallel.all_gather_vectors(local)  # NCCL channel set-up for this message size is not part of the measurement
        barrier()
        ev0.record()
        allv, row0 = parallel.all_gather_vectors(local)
        ev1.record()
        barrier()
        gather_ms = max_over_ranks(ev0.elapsed_time(ev1))
        idx = torch.empty(hi - lo, dtype=torch.int32, device=dev)
        dst = torch.empty(hi - lo, dtype=torch.float32, device=dev)
        barrier()
        ev0.record()
        eng.distance_nearest_device(allv.data_ptr(), nv, row0, hi - lo, idx.data_ptr(), dst.data_ptr(), 0, stream=stream)
        ev1.record()
        barrier()
        near_ms = max_over_ranks(ev0.elapsed_time(ev1))
        # check on a row sample against a brute-force scan in plain float32 torch ops (separate sub / mul / add kernels:
        # the roundings of bl_distance, no FMA): the nearest song's distance bit for bit and the lowest index at it
        ok = True
        probe = torch.arange(0, hi - lo, max(1, (hi - lo) // 64), device=dev)[:64]
        cols = torch.arange(nv, device=dev, dtype=torch.int64)
        for r in probe.tolist():
            dlt = local[r][None, :] - allv
            sq = dlt * dlt
            dd = torch.sqrt(((sq[:, 0] + sq[:, 1]) + sq[:, 2]) + sq[:, 3])
            dd[row0 + r] = float("inf")
            mn = dd.min()
            first_idx = int(torch.where(dd == mn, cols, nv).min())
            ok = ok and float(mn) == float(dst[r]) and first_idx == int(idx[r])
        all_pairs = {"workload": f"BASELINE.json configs[3]: all-pairs bl_distance over {nv} force vectors, fused nearest-"
                                 "neighbour epilogue (matrix never materialised), rows sharded by rank",
                     "n_vectors": nv, "pairs_per_s": float(nv) * nv / (near_ms * 1e-3), "ms": near_ms,
                     "all_gather_ms": gather_ms, "all_gather_bytes": nv * 16, "bruteforce_rows_checked": int(probe.numel()),
                     "bruteforce_argmin_ok": ok}
        del table, allv

    # ---------------- e2e: host buffers through the C-ABI, copies inside the timed region
    Be = min(args.e2e_songs, B)
    try:  # every rank pins its own songs: keep the ranks of one box within half of the host memory that is free
        import psutil
        fit = int(0.5 * psutil.virtual_memory().available / max(world, 1) / (stride * 4))
        Be = max(8, min(Be, fit))
    except Exception:
        pass
    # Ranks of one box do not see the same host link: with all GPUs copying at once, four of the eight GPUs of this pool's
    # boxes get 23 GB/s and four 35 GB/s (alone: 55 GB/s each; tools/h2d_probe.py, profiles/r2_h2d_probe_n8.json). Every rank
    # therefore takes a share of the step's songs in proportion to the copy rate it measures while all ranks copy, so that
    # the ranks finish together; the total stays world x Be songs per step.
    share = Be
    rates = None
    if world > 1:
        probe = torch.empty(64 << 20, dtype=torch.float32, pin_memory=True)  # 256 MiB
        dst = torch.empty_like(probe, device=dev)
        dst.copy_(probe, non_blocking=True)
        barrier()
        ev0.record()
        for _ in range(6):
            dst.copy_(probe, non_blocking=True)
        ev1.record()
        torch.cuda.synchronize()
        mine = 6 * probe.numel() * 4 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
        tr = torch.tensor([mine], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(tr) for _ in range(world)]
        dist.all_gather(allr, tr)
        rates = [float(x.item()) for x in allr]
        share = max(8, min(B, int(round(world * Be * rates[rank] / sum(rates)))))
        del probe, dst
        barrier()
    Be_mine = share
    pinned = torch.empty(Be_mine * stride, dtype=torch.float32, pin_memory=True)
    pinned.copy_(buf[:Be_mine * stride])
    torch.cuda.synchronize()
    ptrs = [pinned.data_ptr() + 4 * (i * stride) for i in range(Be_mine)]
    lens_e = [n_in] * Be_mine
    out_host = np.zeros(Be_mine, dtype=bliss_b200.RESULT_DTYPE)
    for _ in range(max(1, min(args.warmup, 2))):
        eng.analyze_host_ptrs(E.FMT_F32, ptrs, lens_e, out=out_host)
    barrier()
    e2e_steps = max(1, min(args.steps, 8))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.analyze_host_ptrs(E.FMT_F32, ptrs, lens_e, out=out_host)  # synchronous: results are on the host
    torch.cuda.synchronize()
    my_s = time.perf_counter() - t0
    e2e_s = max_over_ranks(my_s)
    if out_host.tobytes() != res[:Be_mine].tobytes():
        for k in out_host.dtype.names:
            d = np.flatnonzero(out_host[k] != res[:Be_mine][k])
            if len(d):
                log(f"  field {k}: {len(d)} songs differ, e.g. song {d[0]}: host {out_host[k][d[0]]!r} device {res[k][d[0]]!r}")
        raise SystemExit("bench.py: host-buffer path and device-resident path disagree")
    songs_all = Be_mine
    if world > 1:
        ts = torch.tensor([float(Be_mine)], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        songs_all = int(ts.item())
    # raw pinned host -> device bandwidth of this box, for context: the e2e path is PCIe-bound
    probe_n = min(Be_mine, 64) * stride
    dprobe = torch.empty(probe_n, dtype=torch.float32, device=dev)
    dprobe.copy_(pinned[:probe_n], non_blocking=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(3):
        dprobe.copy_(pinned[:probe_n], non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    h2d_gbs = 3 * probe_n * 4 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
    del dprobe
    e2e = {"value": songs_all * e2e_steps / e2e_s, "unit": UNIT, "h2d_achieved_gbs": Be_mine * e2e_steps * n_in * 4 / my_s / 1e9,
           "h2d_memcpy_peak_gbs": h2d_gbs, "h2d_bytes_per_step": songs_all * n_in * 4,
           "d2h_bytes_per_step": songs_all * 32, "songs_per_step": songs_all, "steps": e2e_steps,
           "api": "blx_analyze_batch_f32 (include/blx.h), pinned host PCM",
           "rank_shares": None if rates is None else {"rule": "songs per rank in proportion to the host->device rate each rank measures while "
                                                              "all ranks copy at once (the box's host links are not symmetric)",
                                                      "concurrent_h2d_gbs_per_rank": [round(r, 1) for r in rates],
                                                      "rank0_songs": Be_mine, "aggregate_concurrent_h2d_gbs": round(sum(rates), 1)}}

    # ---------------- CPU baseline + parity on a bounded sample of the same songs (rank 0, N = 1 only)
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        S = max(2, min(4 * cores, 128, B))  # ~10-20 s of CPU work on the box's cores
        f32 = np.stack([buf[i * stride:i * stride + n_in].cpu().numpy() for i in range(S)])
        r, _ = cpu_leg_subprocess(f32, 1, 0)
        cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": r["workers"], "kind": r["kind"],
                        "sample": f"first {S} songs of the step's batch, {r['seconds_per_step']:.2f} s wall on {r['workers']} "
                                  f"worker processes ({cores} host cores); front-end excluded; reference src/*.c compiled "
                                  f"verbatim -O3 -std=c99 with a shim FFT in place of fftw3/av_rdft"}
        ref = np.array(r["results"])
        got = np.stack([res[k][:S].astype(np.float64) for k in ("tempo", "amplitude", "frequency", "attack")], axis=1)
        rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-30)
        parity = {"songs": S, "max_rel_err": float(rel.max()), "tolerance": 1e-4, "ok": bool(rel.max() <= 1e-4)}

    # ---------------- the drop-in entry point, file by file (rank 0)
    bl_path = None
    if rank == 0 and not args.no_bl_analyze:
        bl_path = bl_analyze_leg(torch, buf, stride, n_in)

    # ---------------- parity campaign (all ranks) and configs[4] chained
    campaign = None
    if args.parity_songs > 0 and not args.no_cpu:
        t0 = time.perf_counter()
        campaign = parity_campaign(torch, eng, E, dev, stream, rank, world, args.parity_songs, dist)
        campaign["wall_s"] = time.perf_counter() - t0
        if not campaign["ok"]:
            log(f"[rank {rank}] PARITY CAMPAIGN FAILED: {campaign}")
    chain = None
    if args.chain_songs > 0:
        chain = chain_leg(torch, eng, E, dev, stream, rank, world, dist, barrier, max_over_ranks, buf, stride, n_in, B,
                          args.chain_songs, hbm_peak)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (envelope) + f32 (spectrum) + int64 (statistics)", "data": "synthetic",
            "config": workload_config(args, B, world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_kernels": kernels, "spectral_only": spectral, "native_s16": native, "all_pairs": all_pairs, "cpu_baseline": cpu_baseline,
            "parity": campaign if campaign is not None else parity, "parity_sample": parity, "configs4_chained": chain, "bl_analyze_path": bl_path,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--songs-per-step", type=int, default=2048)
    ap.add_argument("--seconds", type=float, default=180.0)
    ap.add_argument("--e2e-songs", type=int, default=256)
    ap.add_argument("--s16-songs", type=int, default=1024)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-spectral", action="store_true")
    ap.add_argument("--no-distance", action="store_true")
    ap.add_argument("--no-bl-analyze", action="store_true")
    ap.add_argument("--distance-vectors", type=int, default=1 << 20)
    ap.add_argument("--parity-songs", type=int, default=4096, help="songs of the parity campaign, all ranks together (0 = off)")
    ap.add_argument("--chain-songs", type=int, default=32768,
                    help="songs per GPU of the configs[4] leg (analysis -> all-gather -> all-pairs); 0 = off")
    ap.add_argument("--_cpu-leg", dest="cpu_leg_path", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--_cpu-energies", dest="cpu_energies", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--_cpu-energy-path", dest="cpu_energy_path", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_leg_path:
        r = cpu_leg(np.load(args.cpu_leg_path, mmap_mode="r"), args.steps, args.warmup, energies=args.cpu_energies,
                    energy_path=args.cpu_energy_path)
        print(json.dumps(r), flush=True)
        return 0
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
