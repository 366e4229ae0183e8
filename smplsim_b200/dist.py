"""Multi-GPU plumbing: one process per GPU, envs sharded by rank, no collective on the physics path.
The only collectives are the PPO-side reductions a learner needs: gradient averaging, advantage moments
(smpl_sim/learning/learning_utils.py:215), RunningNorm batch moments (smpl_sim/learning/running_norm.py:22-29).
Works with backend "nccl" on GPUs and "gloo" on CPU (tests)."""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def shard_range(num_envs_global: int, rank: int, world: int) -> Tuple[int, int]:
    """Static block partition: rank g owns envs [g*N/G, (g+1)*N/G) (remainder spread over the first ranks)."""
    base, rem = divmod(num_envs_global, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_seed(seed: int, rank: int) -> int:
    """Per-rank Philox key so that shards draw independent task targets (same rule as HumanoidBatchB200(rank=...))."""
    return (int(seed) + 0x9E3779B97F4A7C15 * int(rank)) & 0xFFFFFFFFFFFFFFFF


def allreduce_mean_grads(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 8 << 20):
    """Gradient averaging for the policy and value nets: per-layer buckets (last layers first, the order backward produced
    them), each launched asynchronously so that the NVLink transfers of one bucket overlap the packing / unpacking of the next;
    one wait at the end.  (The reference's MLP is 6 layers, 2048...512: ~30 MB of gradients -> 4 buckets.)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in params if p.grad is not None][::-1]
    if not grads:
        return
    world = dist.get_world_size()
    buckets, cur, size = [], [], 0
    for g in grads:
        cur.append(g); size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            buckets.append(cur); cur, size = [], 0
    if cur:
        buckets.append(cur)
    pending = []
    for b in buckets:
        flat = torch.cat([g.reshape(-1) for g in b])
        pending.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, b))
    for work, flat, b in pending:
        work.wait()
        flat /= world
        o = 0
        for g in b:
            n = g.numel()
            g.copy_(flat[o:o + n].view_as(g))
            o += n


def global_moments(x: torch.Tensor, with_count: bool = False):
    """mean, std (population) of x over all ranks -- advantage normalisation (learning_utils.py:215); float64 sums.  With
    ``with_count`` also the global element count (shards may be uneven, see shard_range)."""
    s = torch.stack([x.sum(dtype=torch.float64), (x.double() ** 2).sum(), torch.tensor(float(x.numel()), dtype=torch.float64, device=x.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    mean = s[0] / s[2]
    var = torch.clamp(s[1] / s[2] - mean * mean, min=0.0)
    if with_count:
        return mean.to(x.dtype), var.sqrt().to(x.dtype), float(s[2])
    return mean.to(x.dtype), var.sqrt().to(x.dtype)


def max_over_ranks(v: float, device=None) -> float:
    t = torch.tensor([v], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
