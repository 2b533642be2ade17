"""Env-sharded data parallelism: the only exchange steps of the path.

One process per GPU, `envs_per_rank` environments each, identical parameters on every rank.  K1-K5 need no
communication; the PPO update exchanges (1) the flat gradient buffer once per mini-batch (PPO:244-246: the reference's
losses are `.mean()`s over equally sized shards, so sum / world == gradient of the union batch) and (2) the advantage
statistics (n, sum, sum of squares) once per iteration so that the joint normalisation of RS:149-150 equals the
single-process result.  Everything here works on any device / backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def env_shard(total_envs: int, rank: int, world_size: int):
    """Contiguous shard [start, start + count) of the global env range owned by `rank` (equal shards: weak scaling)."""
    if total_envs % world_size:
        raise ValueError(f"total_envs={total_envs} is not a multiple of world_size={world_size}")
    count = total_envs // world_size
    return rank * count, count


def rank_seed(seed: int, rank: int) -> int:
    """Per-rank seed of the in-kernel Philox stream / of randperm: ranks must draw different randoms (different envs)."""
    return (int(seed) * 1000003 + 7919 * int(rank)) & 0x7FFFFFFF


def grad_scale(world_size: int) -> float:
    """Factor applied inside dwbc_clip_adam_step to the all-reduced (summed) gradient."""
    return 1.0 / world_size


def allreduce_sum_(t: torch.Tensor, world_size: int, group=None) -> torch.Tensor:
    """In-place sum over ranks; a no-op for world_size == 1 (no process group needed)."""
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allreduce_grad_(flat_grad: torch.Tensor, first: int, count: int, world_size: int, group=None) -> torch.Tensor:
    """Sum `flat_grad[first:first+count]` (one contiguous slice of the flat fp32 gradient buffer) over ranks."""
    return allreduce_sum_(flat_grad[first:first + count], world_size, group)


def allreduce_adv_stats_(stats: torch.Tensor, world_size: int, group=None) -> torch.Tensor:
    """`stats` = float64 [3] = (n, sum adv, sum adv^2) of the local, un-normalised advantages."""
    assert stats.dtype == torch.float64 and stats.numel() >= 3
    return allreduce_sum_(stats, world_size, group)


def adv_mean_std(stats: torch.Tensor):
    """(mean, unbiased std) from (n, sum, sum of squares): what dwbc_normalize_advantages applies (RS:150)."""
    n, s, sq = (float(x) for x in stats[:3])
    mean = s / n
    var = max(sq - s * s / n, 0.0) / (n - 1.0)
    return mean, var ** 0.5


def broadcast_params_(flat_params: torch.Tensor, world_size: int, group=None, src: int = 0) -> torch.Tensor:
    """Replicas must start identical (they stay identical because every rank applies the same reduced gradient)."""
    if world_size > 1:
        dist.broadcast(flat_params, src=src, group=group)
    return flat_params


def replicas_identical(flat_params: torch.Tensor, world_size: int, group=None) -> bool:
    """Debug / test helper: bitwise comparison of the parameter replicas."""
    if world_size == 1:
        return True
    lo, hi = flat_params.clone(), flat_params.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return bool(torch.equal(lo, hi))
