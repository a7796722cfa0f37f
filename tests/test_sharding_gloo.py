"""The N>1 path on CPU: pages shard round-robin over ranks with no data-path collective;
only the timing reduction (max over ranks) and the result gather use torch.distributed.
world_size 2, gloo backend, 127.0.0.1 rendezvous."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from page_segmentation_b200.runtime import shard_pages


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_pages, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_pages(n_pages, rank, world)
    # stand-in for the per-rank device work: a deterministic per-page result
    local = torch.tensor([[p, (p * 7919) % 1000] for p in mine], dtype=torch.int64).reshape(-1, 2)
    gathered = [None] * world
    dist.all_gather_object(gathered, local.tolist())
    # timing contract of bench.py: the step time is the max over ranks
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        flat = sorted(x for part in gathered for x in part)
        np.save(os.path.join(out_dir, "gathered.npy"), np.array(flat))
        np.save(os.path.join(out_dir, "tmax.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_sharding_is_a_partition():
    for n, world in [(0, 2), (1, 2), (7, 2), (64, 8), (1024, 8), (5, 8)]:
        parts = [shard_pages(n, r, world) for r in range(world)]
        flat = sorted(p for part in parts for p in part)
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_rank_gloo_page_sharding(tmp_path):
    world, n_pages = 2, 9
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_pages, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    assert got[:, 0].tolist() == list(range(n_pages))                    # every page exactly once, order restorable
    assert got[:, 1].tolist() == [(p * 7919) % 1000 for p in range(n_pages)]
    assert np.load(tmp_path / "tmax.npy")[0] == 11.0
