"""Two-rank NCCL run of the env-sharded update against the single-GPU run on the union batch.  Needs 2 GPUs
(skipped otherwise; the host-side exchange logic is covered on CPU by tests/test_dist_cpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

NL, T, WORLD = 64, 8, 2          # envs per rank, rollout length


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make(n_envs, device, world, group):
    from dwbc_b200.actor_critic import FlatActorCritic
    from dwbc_b200.ppo import FusedPPO
    from test_oracle_golden import ppo_hp
    ac = FlatActorCritic(device=device, seed=0, init_std=[[0.8, 1.0, 1.0] * 4 + [1.0] * 6], num_priv=24, num_hist=10, num_prop=76)
    hp = ppo_hp()
    hp.update(num_learning_epochs=2, num_mini_batches=2)
    alg = FusedPPO(ac, device=device, world_size=world, process_group=group, precision="tf32x3", **hp)
    alg.init_storage(n_envs, T, [860], [None], [18])
    alg.counter = 1500
    return alg


def _worker(rank, port, out):
    import torch.distributed as dist
    from dwbc_b200 import shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=torch.device(dev))
    try:
        r = shard.union_batch_parity(lambda n, w, g: _make(n, dev, w, g), rank, WORLD, dev, NL, T, 2)
        if rank == 0:
            assert r["replicas_identical"], "replicas diverged"
            assert r["advantage_max_abs_diff"] <= 2e-6, r
            # Adam divides by |g| + 1e-8: a reduction-order difference of 1e-9 in a near-zero gradient moves that parameter by up to ~lr,
            # so the criterion is: no entry off by more than lr = 2e-4, and all but 0.1 % of the entries within 5e-6
            assert r["param_max_abs_diff"] < 2e-4 and r["param_frac_beyond_5e_6"] < 1e-3, r
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        out.put((rank, "".join(traceback.format_exception(e))))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_update_equals_union_batch():
    import torch.multiprocessing as mp
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
