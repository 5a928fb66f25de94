"""Multi-GPU host logic: one process per GPU, `torch.distributed` for the plumbing.

The analysis shards by song with no data-path collective. Only an all-pairs distance
request exchanges data: 16 bytes per song are all-gathered (NCCL over NVLink on GPUs, gloo in the CPU
tests) and every rank then computes the row slab of its own songs against all songs.
"""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of rank `rank`: the first n % world ranks get one extra item."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items, world):
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def all_gather_vectors(local, group=None):
    """All-gathers the (n_local, 4) float32 force vectors of every rank, in rank order.

    `local` is a torch tensor on the device the process group communicates on (CUDA for NCCL, CPU for
    gloo). Shards may be ragged: sizes are exchanged first and the payload is padded to the largest
    shard. Returns (all_vectors (n_total, 4), row0) with row0 = index of this rank's first song."""
    import torch
    import torch.distributed as dist
    local = local.reshape(-1, 4).contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local, 0
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    n_max = max(sizes)
    padded = torch.zeros((n_max, 4), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    gathered = torch.empty((world * n_max, 4), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    parts = [gathered[r * n_max:r * n_max + sizes[r]] for r in range(world)]
    return torch.cat(parts, dim=0).contiguous(), sum(sizes[:rank])


def results_to_vectors(results):
    """(n, 4) float32 [tempo, amplitude, frequency, attack] from a RESULT_DTYPE array."""
    return np.stack([results["tempo"], results["amplitude"], results["frequency"], results["attack"]], axis=1).astype(np.float32)


def nearest_neighbours(engine, local_vectors, group=None, with_row_sum=False):
    """All-pairs euclidean distances (bl_distance semantics) of this rank's songs against the songs of
    every rank, reduced on the fly to the nearest other song. `local_vectors`: CUDA tensor (n_local, 4).
    Returns (global index of the nearest song, its distance[, row sum]) as CUDA tensors."""
    import torch
    allv, row0 = all_gather_vectors(local_vectors, group)
    n_rows = local_vectors.reshape(-1, 4).shape[0]
    idx = torch.empty(n_rows, dtype=torch.int32, device=allv.device)
    dst = torch.empty(n_rows, dtype=torch.float32, device=allv.device)
    rsum = torch.empty(n_rows, dtype=torch.float64, device=allv.device) if with_row_sum else None
    st = torch.cuda.current_stream(allv.device).cuda_stream
    engine.distance_nearest_device(allv.data_ptr(), allv.shape[0], row0, n_rows, idx.data_ptr(), dst.data_ptr(),
                                   rsum.data_ptr() if with_row_sum else 0, stream=st)
    return (idx, dst, rsum) if with_row_sum else (idx, dst)


def distance_slab(engine, local_vectors, group=None, cosine=False):
    """Materialises this rank's (n_local, n_total) slab of the distance matrix on its GPU."""
    import torch
    allv, row0 = all_gather_vectors(local_vectors, group)
    n_rows = local_vectors.reshape(-1, 4).shape[0]
    out = torch.empty((n_rows, allv.shape[0]), dtype=torch.float32, device=allv.device)
    st = torch.cuda.current_stream(allv.device).cuda_stream
    engine.distance_rows_device(allv.data_ptr(), allv.shape[0], row0, n_rows, out.data_ptr(), cosine=cosine, stream=st)
    return out
