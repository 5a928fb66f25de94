"""Why does the end-to-end number stop scaling on the 8-GPU box? (per-GPU host->device 53 -> 53 -> 28 -> 23 GB/s
at N = 1 / 2 / 4 / 8.) Launch with torchrun on N ranks: every rank copies from its own pinned buffer to its own GPU
  (a) alone, the other ranks idle, (b) all ranks at once,
for ordinary pinned memory and for write-combined pinned memory, in 256 MiB and 1 GiB pieces. If (a) stays at the
single-GPU rate and (b) falls, the host side of the box (PCIe root complexes / memory controllers shared by several GPUs,
seen from inside a VM with one virtual NUMA node) is the ceiling - not the engine.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py > gpurun_out/h2d_N.json
"""
import ctypes
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nbytes = 2 << 30
    rt = ctypes.CDLL("libcudart.so.12")
    bufs = {}
    bufs["pinned"] = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    p = ctypes.c_void_p()
    if rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04)) == 0:  # cudaHostAllocWriteCombined
        ctypes.memset(p, 1, nbytes)
        bufs["write_combined"] = p
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    bufs["pinned"].fill_(1)
    stream = torch.cuda.Stream(device=dev)
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]

    def copy_rate(kind, piece, reps=3):
        src = bufs[kind].data_ptr() if kind == "pinned" else bufs[kind].value
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            ev0.record()
            for _ in range(reps):
                for off in range(0, nbytes, piece):
                    rt.cudaMemcpyAsync(d.data_ptr() + off, src + off, min(piece, nbytes - off), 1, ctypes.c_void_p(stream.cuda_stream))
            ev1.record()
        stream.synchronize()
        return reps * nbytes / (ev0.elapsed_time(ev1) * 1e-3) / 1e9

    res = {}
    for kind in bufs:
        for piece in (256 << 20, 1 << 30):
            key = f"{kind}_{piece >> 20}MiB"
            copy_rate(kind, piece, 1)
            alone = 0.0
            for turn in range(world):  # (a) one rank at a time
                barrier()
                if turn == rank:
                    alone = copy_rate(kind, piece)
                barrier()
            barrier()
            together = copy_rate(kind, piece)  # (b) all ranks at once
            barrier()
            t = torch.tensor([alone, together], dtype=torch.float64, device=dev)
            if world > 1:
                out = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(out, t)
            else:
                out = [t]
            res[key] = {"alone_gbs": [round(float(o[0]), 1) for o in out], "together_gbs": [round(float(o[1]), 1) for o in out],
                        "together_aggregate_gbs": round(sum(float(o[1]) for o in out), 1)}
    if rank == 0:
        print(json.dumps({"n_gpus": world, "host_cpus": os.cpu_count(), "h2d": res}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
