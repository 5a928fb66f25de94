"""e2e throughput of blx_analyze_batch_f32 vs staging-chunk size (host pinned -> device)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bliss_b200
from bliss_b200 import engine as E
n_in, Be = 7938000, 256
stride = (n_in + 63) // 64 * 64 + 64
pinned = torch.empty(Be * stride, dtype=torch.float32, pin_memory=True)
g = torch.Generator(); g.manual_seed(1)
blk = (torch.rand(stride, generator=g) - 0.5) * 0.4
for i in range(Be):
    pinned[i * stride:(i + 1) * stride] = torch.roll(blk, i * 977)
ptrs = [pinned.data_ptr() + 4 * i * stride for i in range(Be)]
lens = [n_in] * Be
for chunk in [256 << 20, 512 << 20, 1 << 30, 2 << 30, 4 << 30]:
    eng = bliss_b200.Engine(0, chunk_bytes=chunk)
    out = np.zeros(Be, dtype=bliss_b200.RESULT_DTYPE)
    eng.analyze_host_ptrs(E.FMT_F32, ptrs, lens, out=out)
    t0 = time.perf_counter()
    for _ in range(2):
        eng.analyze_host_ptrs(E.FMT_F32, ptrs, lens, out=out)
    dt = (time.perf_counter() - t0) / 2
    print(f"chunk {chunk >> 20} MiB: {Be / dt:.0f} songs/s, {Be * n_in * 4 / dt / 1e9:.1f} GB/s, {dt * 1e3:.1f} ms", flush=True)
    eng.close()
# pure copy baselines over the same pinned buffer: one 8 GB copy, and song-sized pieces
dev = torch.empty(64 * stride, dtype=torch.float32, device="cuda")
for pieces in [1, 64]:
    n = 64 * stride // pieces
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for rep in range(4):
        for k in range(pieces):
            dev[k * n:(k + 1) * n].copy_(pinned[(rep * 64 * stride) + k * n:(rep * 64 * stride) + (k + 1) * n], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"pure H2D, {pieces} pieces per 2 GB: {4 * 64 * stride * 4 / dt / 1e9:.1f} GB/s", flush=True)
