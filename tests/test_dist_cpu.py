"""N>1 host logic on CPU: env sharding, per-rank seeds and the PPO-side reductions over a world_size-2 gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smplsim_b200.dist import allreduce_mean_grads, global_moments, max_over_ranks, rank_seed, shard_range


def test_shard_range_partitions_exactly():
    for n, w in ((65536, 8), (4096, 1), (10, 3), (7, 8)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert len({rank_seed(7, r) for r in range(8)}) == 8


def _worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Linear(4, 3)
        x = torch.full((5, 4), float(rank + 1))
        net(x).sum().backward()
        g_local = net.weight.grad.clone()
        allreduce_mean_grads(net.parameters())
        # grads are linear in the input here: mean over ranks of (rank+1)*5 = 7.5
        assert torch.allclose(net.weight.grad, torch.full_like(g_local, 7.5))
        adv = torch.arange(4, dtype=torch.float32) + 10 * rank          # rank0: 0..3, rank1: 10..13
        mean, std = global_moments(adv)
        full = torch.cat([torch.arange(4.0), torch.arange(4.0) + 10])
        assert abs(mean.item() - full.mean().item()) < 1e-6 and abs(std.item() - full.std(unbiased=False).item()) < 1e-5
        assert max_over_ranks(1.0 + rank) == float(world)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_reductions():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port), nprocs=2, join=True)
