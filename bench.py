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


def cpu_leg(f32, steps, warmup, workers=None):
    """Times the reference's CPU analysers on the songs of `f32` (S x n float32, 44.1 kHz mono).
    The f32 -> int16/22 050 Hz/stereo front-end (oracle/frontend.c, the same arithmetic the GPU fuses
    into pass 1) runs first and is NOT timed: decode/resample is excluded on both sides.
    Worker PROCESSES, one per core: the reference is not thread-safe (fftw planner)."""
    import multiprocessing as mp

    from oracle.binding import REF_SO
    S, n = f32.shape
    cores = os.cpu_count() or 1
    workers = max(1, min(workers or cores, S))
    n16 = 2 * (n // 2)
    shm = mmap.mmap(-1, S * n16 * 2)
    _CPU["f32"] = f32
    _CPU["s16"] = np.frombuffer(shm, dtype=np.int16).reshape(S, n16)
    _CPU["kind"] = "reference" if os.path.exists(REF_SO) else "port"
    _CPU["duration"] = n // RATE_IN
    ctx = mp.get_context("fork")
    times, results = [], None
    with ctx.Pool(workers) as pool:
        pool.map(_cpu_frontend, range(S), chunksize=1)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            results = pool.map(_cpu_analyze, range(S), chunksize=1)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    total = sum(times)
    res = np.zeros((S, 4), dtype=np.float64)
    for i, a, b, c, d in results:
        res[i] = (a, b, c, d)
    return dict(value=S * len(times) / total, seconds_per_step=total / len(times), workers=workers, cores=cores,
                kind=_CPU["kind"], songs=S, results=res.tolist())


def cpu_leg_subprocess(f32, steps, warmup):
    """Runs cpu_leg in a fresh interpreter (no CUDA context to fork)."""
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        path = os.path.join(d, "sample.npy")
        np.save(path, f32)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--_cpu-leg", path, "--steps", str(steps),
                              "--warmup", str(warmup)], check=True, capture_output=True, text=True)
    return json.loads(out.stdout.strip().splitlines()[-1])


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
        "config": workload_config(args, S, 1),
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
    offs = [i * stride for i in range(B)]
    lens = [n_in] * B
    d_out = torch.zeros(B * 8, dtype=torch.int32, device=dev)
    # All timed work and the CUDA events that bracket it go to ONE explicit (non-default) stream: the
    # C-ABI treats a NULL stream as "the engine's own", which torch events on the default stream
    # would not see.
    torch.cuda.synchronize()
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        eng.analyze_device(E.FMT_F32, buf.data_ptr(), offs, lens, d_out.data_ptr(), stream=stream)

    # ---------------- device-resident steps: `value`
    eng.profile(True)
    for _ in range(args.warmup):
        step()
    barrier()
    eng.profile_reset()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
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
    kern_ms_sum = sum(v[0] for v in prof.values()) or 1.0
    kernels = {}
    for name, (ms, n) in prof.items():
        if n == 0:
            continue
        per_launch_ms = ms / n
        ach = alg_bytes.get(name, 0) * B / (per_launch_ms * 1e-3) / 1e9
        kernels[name] = {"share": ms / kern_ms_sum, "ms_per_launch": per_launch_ms, "launches": n,
                         "alg_bytes_per_launch": alg_bytes.get(name, 0) * B, "achieved_gbs": ach, "frac_hbm": ach / hbm_peak}
        if name == "envelope_kernel":
            tf = FP64_FLOP_PER_HOP * hops * B / (per_launch_ms * 1e-3) / 1e12
            kernels[name]["fp64_tflops"] = tf
            kernels[name]["frac_fp64"] = tf / DFMA_PEAK_TFLOPS
    dom = max(kernels, key=lambda k: kernels[k]["share"])
    traffic = None
    try:
        per_song = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
        traffic = per_song * B if per_song else None  # bytes per launch (ncu capture scaled to this batch)
    except Exception:
        pass
    if dom == "envelope_kernel":
        roofline = {"kernel": dom, "bound": "fp64", "achieved": kernels[dom]["fp64_tflops"], "peak": DFMA_PEAK_TFLOPS,
                    "unit": "TFLOP/s", "frac": kernels[dom]["frac_fp64"], "traffic": traffic,
                    "peak_source": "measured DFMA throughput (tools/ubench.cu, profiles/r1_ubench.txt)",
                    "note": "FP64-pipe bound (SURVEY.md §0 F5, §7.3 H3); its HBM fraction is in roofline_kernels"}
    else:
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": hbm_peak,
                    "unit": "GB/s", "frac": kernels[dom]["frac_hbm"], "traffic": traffic, "peak_source": peak_src}

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
            so, sl = offs[:Bs], [n30] * Bs
            for _ in range(max(3, args.warmup)):
                eng.spectral_device(E.FMT_F32, buf.data_ptr(), so, sl, d_freq.data_ptr(), stream=stream)
            eng.profile(True)
            eng.profile_reset()
            barrier()
            reps = 10
            ev0.record()
            for _ in range(reps):
                eng.spectral_device(E.FMT_F32, buf.data_ptr(), so, sl, d_freq.data_ptr(), stream=stream)
            ev1.record()
            barrier()
            sp_ms = max_over_ranks(ev0.elapsed_time(ev1)) / reps
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
        o16, l16 = [i * stride16 for i in range(Bn)], [n16] * Bn
        durs = [int(args.seconds)] * Bn
        d_out16 = torch.zeros(Bn * 8, dtype=torch.int32, device=dev)
        for _ in range(3):
            eng.analyze_device(E.FMT_S16, s16.data_ptr(), o16, l16, d_out16.data_ptr(), durations=durs, stream=stream)
        barrier()
        ev0.record()
        for _ in range(3):
            eng.analyze_device(E.FMT_S16, s16.data_ptr(), o16, l16, d_out16.data_ptr(), durations=durs, stream=stream)
        ev1.record()
        barrier()
        ms16 = max_over_ranks(ev0.elapsed_time(ev1)) / 3
        r16 = np.frombuffer(d_out16.cpu().numpy().tobytes(), dtype=bliss_b200.RESULT_DTYPE)
        native = {"workload": f"{Bn} x {args.seconds:g}-s int16 / 22 050 Hz / stereo songs per GPU (the reference decoder's output "
                              "format), full pipeline", "value": world * Bn / (ms16 * 1e-3), "unit": UNIT, "ms_per_pass": ms16,
                  "bytes_per_song": n16 * 2, "all_status_ok": bool(np.all(r16["status"] == 0))}
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
        # spot check: the reported neighbour really is at the reported distance
        probe = slice(0, min(hi - lo, 4096))
        chk = torch.linalg.vector_norm(local[probe] - allv[idx[probe].long()], dim=1)
        ok = bool(torch.allclose(chk, dst[probe], rtol=1e-5, atol=1e-6))
        all_pairs = {"workload": f"BASELINE.json configs[3]: all-pairs bl_distance over {nv} force vectors, fused nearest-"
                                 "neighbour epilogue (matrix never materialised), rows sharded by rank",
                     "n_vectors": nv, "pairs_per_s": float(nv) * nv / (near_ms * 1e-3), "ms": near_ms,
                     "all_gather_ms": gather_ms, "all_gather_bytes": nv * 16, "spot_check_ok": ok}
        del table, allv

    # ---------------- e2e: host buffers through the C-ABI, copies inside the timed region
    Be = min(args.e2e_songs, B)
    try:  # every rank pins its own songs: keep the ranks of one box within half of the host memory that is free
        import psutil
        fit = int(0.5 * psutil.virtual_memory().available / max(world, 1) / (stride * 4))
        Be = max(8, min(Be, fit))
    except Exception:
        pass
    pinned = torch.empty(Be * stride, dtype=torch.float32, pin_memory=True)
    pinned.copy_(buf[:Be * stride])
    torch.cuda.synchronize()
    ptrs = [pinned.data_ptr() + 4 * o for o in offs[:Be]]
    out_host = np.zeros(Be, dtype=bliss_b200.RESULT_DTYPE)
    for _ in range(max(1, min(args.warmup, 2))):
        eng.analyze_host_ptrs(E.FMT_F32, ptrs, lens[:Be], out=out_host)
    barrier()
    e2e_steps = max(1, min(args.steps, 8))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.analyze_host_ptrs(E.FMT_F32, ptrs, lens[:Be], out=out_host)  # synchronous: results are on the host
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    if out_host.tobytes() != res[:Be].tobytes():
        for k in out_host.dtype.names:
            d = np.flatnonzero(out_host[k] != res[:Be][k])
            if len(d):
                log(f"  field {k}: {len(d)} songs differ, e.g. song {d[0]}: host {out_host[k][d[0]]!r} device {res[k][d[0]]!r}")
        raise SystemExit("bench.py: host-buffer path and device-resident path disagree")
    # raw pinned host -> device bandwidth of this box, for context: the e2e path is PCIe-bound
    probe_n = min(Be, 64) * stride
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
    e2e = {"value": world * Be * e2e_steps / e2e_s, "unit": UNIT, "h2d_achieved_gbs": Be * e2e_steps * n_in * 4 / e2e_s / 1e9,
           "h2d_memcpy_peak_gbs": h2d_gbs, "h2d_bytes_per_step": world * Be * n_in * 4,
           "d2h_bytes_per_step": world * Be * 32, "songs_per_step": world * Be, "steps": e2e_steps,
           "api": "blx_analyze_batch_f32 (include/blx.h), pinned host PCM"}

    # ---------------- CPU baseline + parity on a bounded sample of the same songs (rank 0, N = 1 only)
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        S = max(2, min(4 * cores, 128, B))  # ~10-20 s of CPU work on the box's cores
        f32 = np.stack([buf[i * stride:i * stride + n_in].cpu().numpy() for i in range(S)])
        r = cpu_leg_subprocess(f32, 1, 0)
        cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": r["workers"], "kind": r["kind"],
                        "sample": f"first {S} songs of the step's batch, {r['seconds_per_step']:.2f} s wall on {r['workers']} "
                                  f"worker processes ({cores} host cores); front-end excluded; reference src/*.c compiled "
                                  f"verbatim -O3 -std=c99 with a shim FFT in place of fftw3/av_rdft"}
        ref = np.array(r["results"])
        got = np.stack([res[k][:S].astype(np.float64) for k in ("tempo", "amplitude", "frequency", "attack")], axis=1)
        rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-30)
        parity = {"songs": S, "max_rel_err": float(rel.max()), "tolerance": 1e-4, "ok": bool(rel.max() <= 1e-4)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (envelope) + f32 (spectrum) + int64 (statistics)", "data": "synthetic",
            "config": workload_config(args, B, world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_kernels": kernels, "spectral_only": spectral, "native_s16": native, "all_pairs": all_pairs, "cpu_baseline": cpu_baseline,
            "parity": parity,
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
    ap.add_argument("--s16-songs", type=int, default=256)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-spectral", action="store_true")
    ap.add_argument("--no-distance", action="store_true")
    ap.add_argument("--distance-vectors", type=int, default=1 << 20)
    ap.add_argument("--_cpu-leg", dest="cpu_leg_path", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_leg_path:
        r = cpu_leg(np.load(args.cpu_leg_path, mmap_mode="r"), args.steps, args.warmup)
        print(json.dumps(r), flush=True)
        return 0
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
