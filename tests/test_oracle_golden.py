"""CPU: the oracle (oracle/) reproduces the golden vectors recorded from the unmodified
reference (tests/golden/*.npz, produced by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

import envstate as E
from dwbc_b200 import config as C, synth
from oracle import env_oracle as EO, ppo_oracle as PO

G = os.path.join(os.path.dirname(__file__), "golden")


def test_rand_column_map_agrees():
    for k in ("RAND_GOAL_ORN", "RAND_GOAL_SPH", "RAND_CMD", "RAND_PUSH", "RAND_RST_DOF", "RAND_RST_XY", "RAND_RST_VEL",
              "RAND_RST_CMD", "RAND_RST_GOAL_ORN", "RAND_RST_GOAL_SPH", "RAND_TERRAIN", "RAND_COLS"):
        assert getattr(C, k) == getattr(EO, k)
    assert C.METRIC_NAMES == EO.METRIC_NAMES


@pytest.mark.parametrize("name", ["flat", "full"])
def test_env_oracle_matches_reference_golden(name):
    g = np.load(os.path.join(G, f"env_{name}.npz"))
    N, steps, seed, counter0 = [int(x) for x in g["meta"]]
    p = E.make_params(name, N)
    orc = EO.EnvOracle(p, E.oracle_state(p, E.initial(p, seed)))
    orc.common_step_counter = counter0
    rt = E.runtime(p)
    assert list(g["sum_names"]) == p.sum_slots()
    for t in range(1, steps + 1):
        i = t - 1
        np.testing.assert_array_equal(orc.s.env_origins.numpy(), g["env_origins_pre"][i])
        E.load_sim_into_oracle(orc, p, E.sim_state(p, seed, t, g["env_origins_pre"][i]))
        obs, rew, arew, rst, ex = orc.post_physics_step(torch.from_numpy(synth.rand_table(p, seed, t)), rt)
        np.testing.assert_array_equal(obs[:, :100].numpy(), g["obs100"][i])
        np.testing.assert_array_equal(rew.numpy(), g["rew"][i])
        np.testing.assert_array_equal(arew.numpy(), g["arm_rew"][i])
        np.testing.assert_array_equal(rst.numpy(), g["reset"][i])
        np.testing.assert_array_equal(orc.s.time_out_buf.numpy(), g["time_out"][i])
        np.testing.assert_array_equal(orc.s.commands.numpy(), g["commands"][i])
        np.testing.assert_array_equal(orc.s.ee_goal_sphere.numpy(), g["ee_goal_sphere"][i])
        np.testing.assert_array_equal(orc.s.goal_timer.numpy(), g["goal_timer"][i])
        np.testing.assert_array_equal(orc.s.episode_length_buf.numpy(), g["ep_len"][i])
        if p.measure_heights:
            np.testing.assert_array_equal(orc.measured_heights.numpy(), g["heights"][i])
            np.testing.assert_array_equal(EO.heights_obs(orc.root[:, 2], orc.measured_heights, p.obs_scale_height).numpy(), g["heights_obs"][i])
        if p.terrain_curriculum:                                                          # LR:421-441 (row a21)
            np.testing.assert_array_equal(orc.s.terrain_levels.numpy(), g["terrain_levels"][i])
            np.testing.assert_array_equal(orc.s.env_origins.numpy(), g["env_origins"][i])
        if int(rst.sum()):
            got = np.array([float(ex["episode"][k]) for k in g["stat_names"]], np.float32)
            np.testing.assert_array_equal(got, g["ep_stats"][i])
    np.testing.assert_array_equal(obs.numpy(), g["final_obs"])
    np.testing.assert_array_equal(orc.s.obs_history_buf.numpy(), g["final_hist"])
    np.testing.assert_array_equal(orc.s.root_states_full.numpy(), g["final_root"])
    np.testing.assert_array_equal(orc.s.dof_state.numpy(), g["final_dof"])
    np.testing.assert_array_equal(np.stack([orc.s.episode_sums[k].numpy() for k in g["sum_names"]]), g["final_sums"])
    assert g["reset"].sum() > 0 and g["time_out"].sum() > 0
    if p.terrain_curriculum:      # the fixture really exercises promotions, demotions and the wrap of solved top levels
        lv = np.concatenate([E.initial(p, seed)["terrain_levels"][None], g["terrain_levels"]])
        d = np.diff(lv, axis=0)
        assert (d > 0).sum() >= 20 and (d < 0).sum() >= 20 and (d < -1).sum() >= 1


def ppo_hp():
    """widowGo1 PPO hyper-parameters (WGC:343-366 with RESUME=True, WGC:35)."""
    return dict(value_loss_coef=1.0, use_clipped_value_loss=True, clip_param=0.2, entropy_coef=0.0,
                num_learning_epochs=5, num_mini_batches=4, learning_rate=2e-4, gamma=0.99, lam=0.95,
                max_grad_norm=1.0, min_policy_std=[[0.15, 0.25, 0.25] * 4 + [0.2] * 3 + [0.05] * 3],
                mixing_schedule=[1.0, 0, 1], priv_reg_coef_schedual=[0, 1, 1000, 1000])


def golden_params(g, seed):
    manifest = PO.param_manifest()
    assert [n for n, _ in manifest] == list(g["names"])
    vals = synth.policy_params(manifest, seed)
    init_std = torch.tensor([[0.8, 1.0, 1.0] * 4 + [1.0] * 6])
    return {n: (init_std.clone() if v is None else torch.from_numpy(v).clone()) for (n, _), v in zip(manifest, vals)}


def test_ppo_oracle_matches_reference_golden():
    """BASELINE.json configs[0]: 64 envs x 40 steps, CPU: GAE + losses + post-Adam params."""
    g = np.load(os.path.join(G, "ppo.npz"))
    N, T, seed, counter = [int(x) for x in g["meta"]]
    hp = ppo_hp()
    P = golden_params(g, seed)
    inp = synth.rollout_inputs(N, T, 860, seed)
    obs = torch.from_numpy(inp["obs"])
    actions = torch.from_numpy(g["actions"])
    vals, lps, rews = [], [], []
    for t in range(T):
        mean = PO.actor_mean(P, obs[t])
        eps = (actions[t] - mean) / P["std"]
        a = PO.policy_act(P, obs[t], eps)
        vals.append(a["values"])
        lps.append(PO.log_prob2(a["mean"], P["std"], actions[t]))
        rews.append(PO.bootstrap_rewards(torch.from_numpy(inp["rew"][t]), torch.from_numpy(inp["arm_rew"][t]),
                                         torch.from_numpy(g["values"][t]), torch.from_numpy(inp["time_outs"][t]), hp["gamma"]))
    np.testing.assert_allclose(torch.stack(vals).numpy(), g["values"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(torch.stack(lps).numpy(), g["log_prob"], rtol=0, atol=2e-5)
    np.testing.assert_array_equal(torch.stack(rews).numpy(), g["rewards"])
    dones = torch.from_numpy(inp["dones"]).unsqueeze(-1).to(torch.uint8)
    ret, adv = PO.compute_returns(torch.from_numpy(g["rewards"]), torch.from_numpy(g["values"]), dones,
                                  torch.from_numpy(g["last_values"]), hp["gamma"], hp["lam"])
    np.testing.assert_array_equal(ret.numpy(), g["returns"])
    np.testing.assert_array_equal(adv.numpy(), g["advantages"])
    storage = dict(observations=obs[:T], actions=actions, values=torch.from_numpy(g["values"]), returns=ret,
                   actions_log_prob=torch.from_numpy(g["log_prob"]), advantages=adv)
    snaps = {}

    def record(k, Pn, Gd, when):
        if k == 0 and when == "pre_step":
            snaps["grad1"] = torch.cat([(Gd[n] if Gd[n] is not None else torch.zeros_like(Pn[n])).reshape(-1) for n in Pn])
        if k == 0 and when == "post_step":
            snaps["param1"] = torch.cat([Pn[n].detach().reshape(-1) for n in Pn])

    logs = PO.ppo_update(P, PO.Adam(list(P.keys()), hp["learning_rate"]), storage, torch.from_numpy(g["perm"]).long(), hp,
                         counter, record)
    np.testing.assert_allclose(snaps["grad1"].numpy(), g["grad1"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(snaps["param1"].numpy(), g["param1"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(torch.cat([P[n].reshape(-1) for n in P]).numpy(), g["param20"], rtol=0, atol=3e-6)
    res = g["update_result"]
    assert abs(float(torch.stack([l["value"] for l in logs]).mean()) - res[0]) < 1e-5
    assert abs(float(torch.stack([l["surrogate"] for l in logs]).mean()) - res[1]) < 1e-5
    assert abs(float(torch.stack([l["priv_reg"] for l in logs]).mean()) - res[5]) < 1e-5
    assert logs[0]["mixing_ratio"] == res[3] and logs[0]["priv_reg_coef"] == res[6]


def ts_storage(g, gts, inp, ts):
    """Storage dict of the torque-supervision golden: the rollout of ppo.npz + synth.arm_torque_inputs (make_golden_ts.py)."""
    T = int(gts["meta"][1])
    f = lambda k: torch.from_numpy(g[k])  # noqa: E731
    st = dict(observations=torch.from_numpy(inp["obs"])[:T], actions=f("actions"), values=f("values"), returns=f("returns"),
              actions_log_prob=f("log_prob"), advantages=f("advantages"))
    st.update({k: torch.from_numpy(ts[k]) for k in ("target_arm_torques", "current_arm_dof_pos", "current_arm_dof_vel")})
    return st


def ts_hp(gts, ts):
    hp = ppo_hp()
    hp.update(torque_supervision=True, adaptive_arm_gains=False, torque_supervision_schedule=[float(x) for x in gts["schedule"]],
              arm_coefs=tuple(torch.from_numpy(np.asarray(c, np.float32)) for c in ts["coefs"]))
    return hp


def test_ppo_oracle_torque_supervision_matches_reference_golden():
    """PPO:224-239 (fixed-gain arm model PPO:318-323), switched on: the oracle against the unmodified reference's update() (ppo_ts.npz)."""
    g, gts = np.load(os.path.join(G, "ppo.npz")), np.load(os.path.join(G, "ppo_ts.npz"))
    N, T, seed, counter = [int(x) for x in gts["meta"]]
    P = golden_params(g, seed)
    ts = synth.arm_torque_inputs(N, T, 6, seed)
    hp = ts_hp(gts, ts)
    snaps = {}

    def record(k, Pn, Gd, when):
        if k == 0 and when == "pre_step":
            snaps["grad1"] = torch.cat([(Gd[n] if Gd[n] is not None else torch.zeros_like(Pn[n])).reshape(-1) for n in Pn])

    logs = PO.ppo_update(P, PO.Adam(list(P.keys()), hp["learning_rate"]), ts_storage(g, gts, synth.rollout_inputs(N, T, 860, seed), ts),
                         torch.from_numpy(g["perm"]).long(), hp, counter, record)
    np.testing.assert_allclose(snaps["grad1"].numpy(), gts["grad1"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(torch.cat([P[n].reshape(-1) for n in P]).numpy(), gts["param20"], rtol=0, atol=3e-6)
    res = gts["update_result"]
    np.testing.assert_allclose(np.array([float(l["arm_torques"]) for l in logs]), gts["mb_arm_losses"], rtol=1e-6)
    assert abs(float(np.mean([float(l["arm_torques"]) for l in logs])) - res[2]) < 2e-4 and logs[0]["torque_supervision_weight"] == res[4] == 0.06
    assert abs(float(torch.stack([l["surrogate"] for l in logs]).mean()) - res[1]) < 1e-5


def test_torque_controller_oracle_matches_reference_golden():
    """`_compute_torques` (WG:1262-1295) restatement against the unmodified reference method (tests/golden/make_golden_torques.py)."""
    from dwbc_b200.config import WidowGo1Params
    from oracle import env_oracle as EO
    g = np.load(os.path.join(G, "torques.npz"))
    p = WidowGo1Params(num_envs=int(g["meta"][0]))
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    out = EO.compute_torques(t("actions"), t("dof_pos"), t("dof_vel"), t("motor"), torch.tensor(p.p_gains), torch.tensor(p.d_gains),
                             torch.tensor(p.action_scale), torch.tensor(p.default_dof_pos), torch.tensor(p.torque_limits))
    np.testing.assert_array_equal(out.numpy(), g["torques"])
    assert np.all(g["torques"][:, p.num_actions:] == 0.0)
