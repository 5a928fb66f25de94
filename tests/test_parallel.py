"""CPU: the multi-GPU host logic on gloo with world size 2 (sharding, ragged all-gather of force
vectors). The distance kernels themselves are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest

from bliss_b200 import parallel


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = parallel.shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)
        allv = rng.standard_normal((n_total, 4)).astype(np.float32)
        lo, hi = parallel.shard_range(n_total, rank, world)
        # make the shards ragged on purpose: rank 0 gives three of its songs to nobody (drops them)
        if rank == 0:
            hi -= 3
        got, row0 = parallel.all_gather_vectors(torch.from_numpy(allv[lo:hi].copy()))
        np.save(os.path.join(out_dir, f"r{rank}.npy"), got.numpy())
        np.save(os.path.join(out_dir, f"row0_{rank}.npy"), np.array([row0, lo, hi]))
    finally:
        dist.destroy_process_group()


def test_ragged_all_gather_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    n_total, world = 37, 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(123)
    allv = rng.standard_normal((n_total, 4)).astype(np.float32)
    lo0, hi0 = parallel.shard_range(n_total, 0, world)
    lo1, hi1 = parallel.shard_range(n_total, 1, world)
    expect = np.concatenate([allv[lo0:hi0 - 3], allv[lo1:hi1]])
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npy")
        assert np.array_equal(got, expect)
    assert np.load(tmp_path / "row0_0.npy")[0] == 0
    assert np.load(tmp_path / "row0_1.npy")[0] == hi0 - 3 - lo0


def test_results_to_vectors():
    import bliss_b200
    r = np.zeros(3, dtype=bliss_b200.RESULT_DTYPE)
    r["tempo"], r["amplitude"], r["frequency"], r["attack"] = [1, 2, 3], [4, 5, 6], [7, 8, 9], [10, 11, 12]
    v = parallel.results_to_vectors(r)
    assert v.dtype == np.float32 and v.shape == (3, 4) and v[1].tolist() == [2, 5, 8, 11]
