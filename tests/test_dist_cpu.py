"""world_size-2 gloo tests (CPU) of the N>1 path's host logic: env sharding, gradient all-reduce + 1/W scaling,
advantage-statistics all-reduce.  The per-rank math is done by the oracle (this is a test), the exchange by the
product's `shard` module; the result must equal the single-process result on the union batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

N, T, WORLD = 32, 8, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup():
    from dwbc_b200 import synth
    from oracle import ppo_oracle as PO
    from test_oracle_golden import ppo_hp
    manifest = PO.param_manifest()
    vals = synth.policy_params(manifest, 3)
    init_std = torch.tensor([[0.8, 1.0, 1.0] * 4 + [1.0] * 6])
    P = {n: (init_std.clone() if v is None else torch.from_numpy(v).clone()) for (n, _), v in zip(manifest, vals)}
    inp = synth.rollout_inputs(N, T, 860, 5)
    g = torch.Generator().manual_seed(0)
    full = dict(observations=torch.from_numpy(inp["obs"])[:T].contiguous(), actions=torch.randn(T, N, 18, generator=g),
                values=torch.randn(T, N, 2, generator=g), rewards=torch.randn(T, N, 2, generator=g),
                actions_log_prob=torch.randn(T, N, 2, generator=g) - 20.0, last_values=torch.randn(N, 2, generator=g),
                dones=torch.from_numpy(inp["dones"]).unsqueeze(-1).to(torch.uint8))
    return PO, P, full, ppo_hp()


def _raw_advantages(PO, st, hp):
    T_ = st["rewards"].shape[0]
    returns = torch.zeros_like(st["values"])
    adv = 0
    for t in reversed(range(T_)):                       # RS:141-147 without the normalisation of RS:149-150
        nxt = st["last_values"] if t == T_ - 1 else st["values"][t + 1]
        m = 1.0 - st["dones"][t].float()
        delta = st["rewards"][t] + m * hp["gamma"] * nxt - st["values"][t]
        adv = delta + m * hp["gamma"] * hp["lam"] * adv
        returns[t] = adv + st["values"][t]
    return returns, returns - st["values"]


def _flat(G, names, P):
    return torch.cat([(G[n] if G[n] is not None else torch.zeros_like(P[n])).reshape(-1) for n in names])


def _grad(PO, P, mb, hp, counter):
    names = list(P.keys())
    Q = {n: P[n].clone().requires_grad_(True) for n in names}
    loss, _ = PO.minibatch_loss(Q, mb, hp, counter)
    loss.backward()
    return _flat({n: (Q[n].grad if Q[n].grad is not None else None) for n in names}, names, P)


def _worker(rank, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        torch.set_num_threads(1)
        from dwbc_b200 import shard
        PO, P, full, hp = _setup()
        start, count = shard.env_shard(N, rank, WORLD)
        loc = {k: (v[start:start + count] if k == "last_values" else v[:, start:start + count]) for k, v in full.items()}
        # ---- (2) advantage statistics: local raw advantages -> (n, sum, sumsq) -> all-reduce -> normalise
        ret, adv = _raw_advantages(PO, loc, hp)
        a64 = adv.double()
        stats = torch.tensor([a64.numel(), a64.sum(), (a64 * a64).sum()], dtype=torch.float64)
        shard.allreduce_adv_stats_(stats, WORLD)
        mean, std = shard.adv_mean_std(stats)
        adv_n = ((adv.double() - mean) / (std + 1e-8)).float()
        ret_u, adv_u = PO.compute_returns(full["rewards"], full["values"], full["dones"], full["last_values"], hp["gamma"], hp["lam"])
        assert torch.equal(ret, ret_u[:, start:start + count])
        np.testing.assert_allclose(adv_n.numpy(), adv_u[:, start:start + count].numpy(), rtol=0, atol=2e-6)
        # ---- (1) gradient: per-rank mini-batch = this rank's rows of the union mini-batch
        loc.update(returns=ret, advantages=adv_u[:, start:start + count].contiguous())
        fullst = dict(full, returns=ret_u, advantages=adv_u)
        idx_local = torch.randperm(T * count, generator=torch.Generator().manual_seed(11))[: T * count // 2]
        t_, e_ = idx_local // count, idx_local % count
        g_local = _grad(PO, P, PO.gather(loc, idx_local), hp, 1500)
        flat = g_local.clone()
        shard.allreduce_grad_(flat, 0, flat.numel(), WORLD)
        flat *= shard.grad_scale(WORLD)
        # union mini-batch: both ranks' rows (each rank draws the same local permutation here, so the union is known everywhere)
        idx_union = torch.cat([t_ * N + (r * count + e_) for r in range(WORLD)])
        g_union = _grad(PO, P, PO.gather(fullst, idx_union), hp, 1500)
        np.testing.assert_allclose(flat.numpy(), g_union.numpy(), rtol=0, atol=2e-6 * float(g_union.abs().max()) + 1e-9)
        # slice all-reduce only touches the slice (update_dagger reduces the history-encoder slice only)
        buf = torch.full((64,), float(rank + 1))
        shard.allreduce_grad_(buf, 16, 8, WORLD)
        assert torch.equal(buf[16:24], torch.full((8,), 3.0)) and torch.equal(buf[:16], torch.full((16,), float(rank + 1)))
        # ---- replicas: broadcast then identical reduced gradient -> identical parameters
        names = list(P.keys())
        fp = torch.cat([P[n].reshape(-1) for n in names]) + (0.01 * rank)
        shard.broadcast_params_(fp, WORLD)
        assert shard.replicas_identical(fp, WORLD)
        fp2 = fp - 2e-4 * flat[: fp.numel()]
        assert shard.replicas_identical(fp2, WORLD)
        assert not shard.replicas_identical(fp2 + rank, WORLD)
        assert shard.rank_seed(1, 0) != shard.rank_seed(1, 1)
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        out.put((rank, "".join(traceback.format_exception(e))))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_exchange_matches_union_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, out)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=280) for _ in procs)
    for p in procs:
        p.join(30)
    assert res == {0: "ok", 1: "ok"}, res


def test_env_shard_and_world1_noops():
    from dwbc_b200 import shard
    assert shard.env_shard(32768, 3, 8) == (12288, 4096)
    with pytest.raises(ValueError):
        shard.env_shard(100, 0, 8)
    t = torch.ones(4)
    assert shard.allreduce_sum_(t, 1) is t and shard.grad_scale(1) == 1.0 and shard.replicas_identical(t, 1)
    m, s = shard.adv_mean_std(torch.tensor([4.0, 10.0, 30.0], dtype=torch.float64))
    x = torch.tensor([1.0, 2.0, 3.0, 4.0], dtype=torch.float64)
    assert abs(m - float(x.mean())) < 1e-12 and abs(s - float(x.std())) < 1e-12
