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


# ----------------------------------------------------------------------------------------------------------------------
# Equivalence of the env-sharded update with the single-process update on the union batch ("8-GPU run on 8 shards ==
# 1-GPU run on the concatenated storage with the same per-shard permutations").  Used by tests/test_gpu_dist.py and by the
# preamble of `bench.py --gpus N` (its `dist_parity` block), so that the driver's own multi-GPU run carries the proof.
# ----------------------------------------------------------------------------------------------------------------------
def synthetic_rollout(total_envs: int, T: int, seed: int = 3):
    """Seeded rollout data of the whole (union) batch, identical on every rank (CPU generator)."""
    g = torch.Generator().manual_seed(seed)
    N = total_envs
    return dict(obs=torch.randn(T + 1, N, 860, generator=g), actions=torch.randn(T, N, 18, generator=g), values=torch.randn(T, N, 2, generator=g),
                rewards=torch.randn(T, N, 2, generator=g), log_prob=torch.randn(T, N, 2, generator=g) - 20.0,
                dones=(torch.rand(T, N, 1, generator=g) < 0.05).to(torch.uint8))


def _fill_storage(alg, d, lo, hi, device, T):
    s = alg.storage
    s._obs_all.copy_(d["obs"][:, lo:hi].to(device))
    for k, src in (("actions", "actions"), ("values", "values"), ("rewards", "rewards"), ("actions_log_prob", "log_prob"), ("dones", "dones")):
        getattr(s, k).copy_(d[src][:, lo:hi].to(device))
    s.step = T


def union_batch_parity(make_alg, rank: int, world_size: int, device, envs_per_rank: int = 64, T: int = 8, mini_batches: int = 2, group=None):
    """Runs `update()` on this rank's shard (gradient and advantage statistics all-reduced over `world_size` ranks) and, on rank 0, the same
    update in ONE process on the union batch, whose mini-batch k is the concatenation of every rank's mini-batch k.

    make_alg(n_envs, world_size, group) -> FusedPPO with storage for (n_envs, T) and `mini_batches` mini-batches.
    Returns (on rank 0; None elsewhere) the measured differences: Adam divides by |g| + 1e-8, so a reduction-order difference of 1e-9 in a
    near-zero gradient moves that entry by up to ~lr; the statistics are the max and the fraction of entries beyond 5e-6."""
    NL, W = envs_per_rank, world_size
    d = synthetic_rollout(NL * W, T)
    alg = make_alg(NL, W, group)
    _fill_storage(alg, d, rank * NL, (rank + 1) * NL, device, T)
    alg.compute_returns(d["obs"][T, rank * NL:(rank + 1) * NL].to(device))
    perm = torch.randperm(T * NL, generator=torch.Generator().manual_seed(11))          # the same local permutation on every rank
    adv = alg.storage.advantages.clone()
    alg.update(indices=perm.to(device))
    flat = alg.actor_critic.flat.clone()
    identical = replicas_identical(flat, W, group)
    if rank != 0:
        return None
    N = NL * W
    ref = make_alg(N, 1, None)
    _fill_storage(ref, d, 0, N, device, T)
    ref.compute_returns(d["obs"][T].to(device))
    adv_diff = float((adv - ref.storage.advantages[:, :NL]).abs().max())
    mbs = perm.numel() // mini_batches
    t_, e_ = perm // NL, perm % NL
    union = torch.cat([torch.cat([t_[k * mbs:(k + 1) * mbs] * N + r * NL + e_[k * mbs:(k + 1) * mbs] for r in range(W)]) for k in range(mini_batches)])
    ref.update(indices=union.to(device))
    diff = (ref.actor_critic.flat - flat).abs()
    return dict(world_size=W, envs_per_rank=NL, steps=T, mini_batches=mini_batches, replicas_identical=bool(identical),
                advantage_max_abs_diff=adv_diff, param_max_abs_diff=float(diff.max()), param_frac_beyond_5e_6=float((diff > 5e-6).float().mean()),
                lr=float(ref.learning_rate))
