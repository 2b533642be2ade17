"""Fused PPO with the reference's method surface (`rsl_rl.algorithms.PPO`,
rsl_rl/rsl_rl/algorithms/ppo.py:38-325, cited PPO:line): `act`, `process_env_step`,
`compute_returns`, `update` (7-tuple), `update_dagger` (float), `enforce_min_std`,
attributes `actor_critic`, `storage`, `learning_rate`, `optimizer`, `counter`.

Every numerical step runs in libdwbc kernels; this class only sequences launches:
  act            -> dwbc_policy_act, writing straight into the storage rows (no RS:95-114 copies)
  process_env_step -> dwbc_store_rewards (time-out bootstrap PPO:133-134)
  compute_returns  -> dwbc_critic_values + dwbc_gae
  update         -> per mini-batch: dwbc_ppo_minibatch_grad [+ NCCL all-reduce] + dwbc_clip_adam_step
  update_dagger  -> per mini-batch: dwbc_dagger_minibatch_grad [+ all-reduce] + dwbc_clip_adam_step

Host<->device synchronisation happens once per update (to return the mean losses), not three
times per mini-batch as in PPO:248-250.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from . import shard
from .actor_critic import FlatActorCritic
from .storage import FusedRolloutStorage


class _AdamState:
    """Flat Adam state of one parameter group; `state_dict()` mimics torch.optim.Adam's layout."""

    def __init__(self, ac: FlatActorCritic, first, count, lr):
        self.ac, self.first, self.count, self.lr = ac, first, count, lr
        self.m = torch.zeros_like(ac.flat)
        self.v = torch.zeros_like(ac.flat)
        self.step = 0
        self.param_groups = [dict(lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0)]

    def state_dict(self):
        names = [n for n in self.ac.offsets if self.first <= self.ac.offsets[n] < self.first + self.count]
        um, uv = self.ac.unflat(self.m), self.ac.unflat(self.v)
        state = {i: dict(step=torch.tensor(float(self.step)), exp_avg=um[n].clone(), exp_avg_sq=uv[n].clone())
                 for i, n in enumerate(names)} if self.step > 0 else {}
        return dict(state=state, param_groups=[dict(self.param_groups[0], params=list(range(len(names))))])

    def load_state_dict(self, sd):
        names = [n for n in self.ac.offsets if self.first <= self.ac.offsets[n] < self.first + self.count]
        um, uv = self.ac.unflat(self.m), self.ac.unflat(self.v)
        for i, n in enumerate(names):
            if i in sd["state"]:
                um[n].copy_(sd["state"][i]["exp_avg"])
                uv[n].copy_(sd["state"][i]["exp_avg_sq"])
                self.step = int(sd["state"][i]["step"])


class FusedPPO:
    def __init__(self, actor_critic: FlatActorCritic, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998,
                 lam=0.95, value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, max_grad_norm=1.0,
                 use_clipped_value_loss=True, schedule="fixed", desired_kl=0.01, device="cuda:0",
                 mixing_schedule=(0.5, 2000, 4000), torque_supervision=False, torque_supervision_schedule=(0.1, 1000, 1000),
                 adaptive_arm_gains=False, min_policy_std=None, dagger_update_freq=20, priv_reg_coef_schedual=(0, 0, 0, 1),
                 world_size=1, process_group=None, precision="tf32x3"):
        if adaptive_arm_gains:
            raise L.DwbcError("adaptive_arm_gains (a 12-output arm head, AC:111-125,214-215; off for widowGo1, WGC:168) is not implemented")
        if schedule != "fixed":
            raise L.DwbcError("only schedule='fixed' (WGC:352) is implemented")
        self.device = torch.device(device)
        self.actor_critic = actor_critic
        self.storage = None
        self.learning_rate, self.schedule, self.desired_kl = learning_rate, schedule, desired_kl
        self.clip_param, self.num_learning_epochs, self.num_mini_batches = clip_param, num_learning_epochs, num_mini_batches
        self.value_loss_coef, self.entropy_coef, self.gamma, self.lam = value_loss_coef, entropy_coef, gamma, lam
        self.max_grad_norm, self.use_clipped_value_loss = max_grad_norm, use_clipped_value_loss
        self.min_policy_std = None if min_policy_std is None else torch.tensor(min_policy_std, device=self.device, dtype=torch.float).reshape(-1)
        self.mixing_schedule, self.priv_reg_coef_schedual = list(mixing_schedule), list(priv_reg_coef_schedual)
        # arm torque supervision with fixed gains (PPO:224-239, 318-323): off in the shipped config (WGC:173), on in PPO's own defaults (PPO:57)
        self.torque_supervision, self.adaptive_arm_gains = bool(torque_supervision), False
        self.torque_supervision_schedule = list(torque_supervision_schedule)
        self._arm_coefs = None
        self.dagger_update_freq = dagger_update_freq
        self.counter = 0
        self.world_size, self.process_group = world_size, process_group
        self._zh_all = None
        self._packed = False          # tensor-core weight images in the workspace match the current parameters (rollout reuse)
        self._eps_all, self._eps_valid = None, False
        self.precision = precision
        ac = actor_critic
        self.optimizer = _AdamState(ac, 0, ac.num_params, learning_rate)                  # PPO:75
        hf, hc = ac.hist_range
        self.hist_encoder_optimizer = _AdamState(ac, hf, hc, learning_rate)               # PPO:79
        self.grad = torch.zeros_like(ac.flat)
        self._losses = torch.zeros(5, device=self.device)        # surrogate, value, priv_reg, entropy, arm torques
        self._norm_scratch = torch.zeros(2, dtype=torch.float64, device=self.device)
        self._grad_norm = torch.zeros(1, device=self.device)
        self._ws = None
        self._ws_rows = 0
        self._hp = L.PpoHyper()
        self._lib = L.lib()
        self.transition = FusedRolloutStorage.Transition()
        self._eps = None
        self.generator = None

    # ------------------------------------------------------------------ plumbing
    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        self.storage = FusedRolloutStorage(num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape, self.device)
        self._eps = torch.zeros(num_envs, *action_shape, device=self.device)
        self._last_values = torch.zeros(num_envs, 2, device=self.device)
        self._act_tmp = [torch.zeros(num_envs, *action_shape, device=self.device) for _ in range(3)] + \
                        [torch.zeros(num_envs, 2, device=self.device) for _ in range(2)]
        self._workspace(max(num_envs, num_envs * num_transitions_per_env // self.num_mini_batches))
        if self.torque_supervision:
            self.storage.enable_torque_supervision(self.actor_critic.num_arm_actions)

    @property
    def precision(self):
        return self._precision

    @precision.setter
    def precision(self, value):
        """'fp32' (CUDA-core GEMMs, parity anchor), 'tf32' (tcgen05, truncated 10-bit-mantissa operands) or 'tf32x3' (tcgen05,
        error-compensated three-product split: fp32-grade).  Travels to the kernels per call inside DwbcNetCfg.precision."""
        if value not in L.PRECISIONS:
            raise L.DwbcError("precision must be one of " + ", ".join(L.PRECISIONS))
        self._precision = value
        self.actor_critic.net_cfg.precision = L.PRECISIONS[value]
        self._packed = False

    def _set_precision(self):
        self.actor_critic.net_cfg.precision = L.PRECISIONS[self._precision]

    def params_changed(self):
        """Call after writing `actor_critic.flat` from outside (load_state_dict does it): the cached weight images are stale."""
        self._packed = False

    def _workspace(self, rows):
        if self._ws is None or rows > self._ws_rows:
            nbytes = self._lib.dwbc_workspace_bytes(C.addressof(self.actor_critic.net_cfg), rows)
            if nbytes < 0:
                raise L.DwbcError("dwbc_workspace_bytes rejected the network configuration")
            self._ws = torch.zeros(nbytes // 4 + 64, device=self.device)
            self._ws_rows = rows
        return self._ws

    def test_mode(self):
        pass

    def train_mode(self):
        pass

    def set_arm_default_coeffs(self, default_arm_p_gains, default_arm_d_gains, default_arm_dof_pos):
        """PPO:307-310 (called at OPR:91).  Kept as one device tensor [3, n_arm] = p gains, d gains, default positions, each broadcast to
        the arm joints the way PPO:318-323 broadcasts them against the [M, n_arm] batch."""
        self.default_arm_p_gains, self.default_arm_d_gains, self.default_arm_dof_pos = default_arm_p_gains, default_arm_d_gains, default_arm_dof_pos
        if not self.torque_supervision:
            return
        n_arm = self.actor_critic.num_arm_actions
        rows = []
        for x in (default_arm_p_gains, default_arm_d_gains, default_arm_dof_pos):
            x = torch.as_tensor(x, dtype=torch.float, device=self.device)
            if x.dim() > 1 and x.shape[0] != 1:
                raise L.DwbcError("arm coefficients must broadcast to [n_arm] (per-env coefficients are not supported)")
            x = x.reshape(-1) if x.dim() else x
            if x.dim() and x.numel() not in (1, n_arm):
                raise L.DwbcError(f"arm coefficient with {x.numel()} entries does not broadcast to the {n_arm} arm joints (PPO:318-323)")
            rows.append(torch.broadcast_to(x, (n_arm,)))
        self._arm_coefs = torch.stack(rows).contiguous()

    def get_torque_supervision_weight(self):
        sch = self.torque_supervision_schedule
        return (1 - min(max((self.counter - sch[1]) / sch[2], 0), 1)) * sch[0]                   # PPO:304-305

    # ------------------------------------------------------------------ rollout
    def act(self, obs, critic_obs=None, hist_encoding=False, eps=None):
        """PPO:115-127.  Outputs land directly in storage row `storage.step` when a storage exists."""
        ac, s = self.actor_critic, self.storage
        self._set_precision()
        n = obs.shape[0]
        if eps is None:
            na = ac.num_leg_actions + ac.num_arm_actions
            if s is not None and n == s.num_envs and s.step < s.num_transitions_per_env:
                # the standard normals of a whole rollout are drawn by ONE generator launch at its first step (Normal.sample() of AC:337-339
                # draws the same distribution once per step)
                if self._eps_all is None or self._eps_all.shape[:2] != (s.num_transitions_per_env, n):
                    self._eps_all = torch.empty(s.num_transitions_per_env, n, na, device=self.device)
                    self._eps_valid = False
                if s.step == 0 or not self._eps_valid:
                    self._eps_all.normal_(generator=self.generator)
                    self._eps_valid = True
                eps = self._eps_all[s.step]
            else:
                if self._eps is None or self._eps.shape[0] != n:
                    self._eps = torch.empty(n, na, device=self.device)
                eps = self._eps.normal_(generator=self.generator)
        if s is not None and s.step < s.num_transitions_per_env and n == s.num_envs:
            t = s.step
            if obs.data_ptr() != s.observations[t].data_ptr():
                s.observations[t].copy_(obs)                                              # RS:98
            acts, vals, lp, mu, sg = s.actions[t], s.values[t], s.actions_log_prob[t], s.mu[t], s.sigma[t]
            obs_c = s.observations[t]
        else:
            acts, mu, sg, vals, lp = self._act_tmp if n == self._act_tmp[0].shape[0] else \
                [torch.zeros(n, ac.num_leg_actions + ac.num_arm_actions, device=self.device) for _ in range(3)] + \
                [torch.zeros(n, 2, device=self.device) for _ in range(2)]
            obs_c = obs.contiguous()
        ws = self._workspace(n)
        key = (n, int(bool(hist_encoding)), ac.flat._version)
        packed = self._packed and self._packed_key == key
        L.check(self._lib.dwbc_policy_act(C.addressof(ac.net_cfg), L.ptr(ac.flat), L.ptr(obs_c), obs_c.stride(0), L.ptr(eps),
                                          int(bool(hist_encoding)), L.ptr(acts), L.ptr(vals), L.ptr(lp), L.ptr(mu), L.ptr(sg), n,
                                          int(packed), L.ptr(ws), L.stream_ptr()), "dwbc_policy_act")
        self._packed, self._packed_key = True, key
        tr = self.transition
        tr.actions, tr.values, tr.actions_log_prob, tr.action_mean, tr.action_sigma = acts, vals, lp, mu, sg
        tr.observations = tr.critic_observations = obs
        return acts

    def process_env_step(self, rewards, arm_rewards, dones, infos):
        """PPO:129-146 + RS:95-114 (the other transition fields were already written by `act`)."""
        s = self.storage
        if s.step >= s.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")                               # RS:96-97
        t = s.step
        stored = infos.get("dwbc_stored_rows") if isinstance(infos, dict) else None
        if stored is None or stored != (s.rewards[t].data_ptr(), s.dones[t].data_ptr()):
            # (else: the post-physics kernel already wrote this step's rewards / dones rows, FusedWidowGo1Core.set_transition_target)
            to = infos.get("time_outs") if isinstance(infos, dict) else None
            to8 = None if to is None else (to if to.dtype in (torch.uint8, torch.bool) else (to != 0)).contiguous()
            d8 = (dones if dones.dtype in (torch.uint8, torch.bool) else (dones != 0)).contiguous()
            L.check(self._lib.dwbc_store_rewards(L.ptr(rewards.float().contiguous(), torch.float32), L.ptr(arm_rewards.float().contiguous(), torch.float32),
                                                 L.ptr(s.values[t]), L.ptr(to8, (torch.uint8, torch.bool)), L.ptr(d8, (torch.uint8, torch.bool)), self.gamma,
                                                 L.ptr(s.rewards[t]), L.ptr(s.dones[t]), s.num_envs, L.stream_ptr()), "dwbc_store_rewards")
        if self.torque_supervision and isinstance(infos, dict) and "target_arm_torques" in infos:          # PPO:136-142, RS:108-111
            s.target_arm_torques[t].copy_(infos["target_arm_torques"])
            s.current_arm_dof_pos[t].copy_(infos["current_arm_dof_pos"])
            s.current_arm_dof_vel[t].copy_(infos["current_arm_dof_vel"])
        s.step += 1
        self.transition.clear()

    def compute_returns(self, last_critic_obs):
        """PPO:148-150."""
        ac, s = self.actor_critic, self.storage
        self._set_precision()
        obs = last_critic_obs.contiguous()
        L.check(self._lib.dwbc_critic_values(C.addressof(ac.net_cfg), L.ptr(ac.flat), L.ptr(obs), obs.stride(0),
                                             L.ptr(self._last_values), obs.shape[0], L.ptr(self._workspace(obs.shape[0])),
                                             L.stream_ptr()), "dwbc_critic_values")
        self._packed = False                      # dwbc_critic_values re-packs the critic into the same workspace region
        s.compute_returns(self._last_values, self.gamma, self.lam, self.world_size, self.process_group)

    # ------------------------------------------------------------------ schedules (PPO:178-179, 301-302)
    def get_value_mixing_ratio(self):
        return min(max((self.counter - self.mixing_schedule[1]) / self.mixing_schedule[2], 0), 1) * self.mixing_schedule[0]

    def get_priv_reg_coef(self):
        sch = self.priv_reg_coef_schedual
        stage = min(max((self.counter - sch[2]), 0) / sch[3], 1)
        return stage * (sch[1] - sch[0]) + sch[0]

    def _fill_hp(self):
        h = self._hp
        h.clip_param, h.value_loss_coef, h.entropy_coef = self.clip_param, self.value_loss_coef, self.entropy_coef
        h.priv_reg_coef, h.mixing_ratio = self.get_priv_reg_coef(), self.get_value_mixing_ratio()
        h.use_clipped_value_loss = int(self.use_clipped_value_loss)
        h.max_grad_norm, h.lr, h.beta1, h.beta2, h.adam_eps = self.max_grad_norm, self.learning_rate, 0.9, 0.999, 1e-8
        h.grad_scale = shard.grad_scale(self.world_size)
        if self.torque_supervision:
            if self._arm_coefs is None:
                raise L.DwbcError("torque_supervision needs set_arm_default_coeffs() first (OPR:91)")
            h.torque_supervision_weight, h.arm_coefs = self.get_torque_supervision_weight(), self._arm_coefs.data_ptr()
        else:
            h.torque_supervision_weight, h.arm_coefs = 0.0, None
        return h

    def _allreduce(self, first, count):
        shard.allreduce_grad_(self.grad, first, count, self.world_size, self.process_group)

    # ------------------------------------------------------------------ update (PPO:152-263)
    def update(self, indices=None, on_step=None):
        ac, s, hp = self.actor_critic, self.storage, self._fill_hp()
        self._set_precision()
        self._packed = False                      # the parameters move (and the workspace is re-used with another row count)
        if indices is None:
            indices, _ = s.draw_indices(self.num_mini_batches, self.generator)
        indices = indices.to(torch.int64).contiguous()
        mbs = indices.numel() // self.num_mini_batches
        ws = self._workspace(mbs)
        self._losses.zero_()
        k = 0
        # The regulariser target z_hist (PPO:175-176) is detached, so update() never moves the history encoder (zero gradient ->
        # zero Adam step): evaluate it once per storage row instead of once per (epoch, row).
        total = s.num_envs * s.num_transitions_per_env
        lld = (ac.priv_dims[-1] + 3) // 4 * 4
        if self._zh_all is None or self._zh_all.shape[0] != total:
            self._zh_all = torch.zeros(total, lld, device=self.device)
        obs_flat = s.observations.view(total, -1)
        for r0 in range(0, total, mbs):
            nrow = min(mbs, total - r0)
            L.check(self._lib.dwbc_hist_latent(C.addressof(ac.net_cfg), L.ptr(ac.flat), L.ptr(obs_flat[r0:]), obs_flat.stride(0),
                                               L.ptr(self._zh_all[r0:]), lld, nrow, L.ptr(ws), L.stream_ptr()), "dwbc_hist_latent")
        s.set_hist_latent(self._zh_all)
        for batch_idx in s.mini_batch_generator(self.num_mini_batches, self.num_learning_epochs, indices):
            L.check(self._lib.dwbc_ppo_minibatch_grad(C.addressof(ac.net_cfg), L.ptr(ac.flat), s.c_struct_ptr(), L.ptr(batch_idx), mbs,
                                                      C.addressof(hp), L.ptr(self.grad), L.ptr(self._losses), L.ptr(ws), L.stream_ptr()),
                    "dwbc_ppo_minibatch_grad")
            self._allreduce(0, ac.num_params)
            if on_step is not None:
                on_step(k, "grad")
            self.optimizer.step += 1
            L.check(self._lib.dwbc_clip_adam_step(L.ptr(ac.flat), L.ptr(self.grad), L.ptr(self.optimizer.m), L.ptr(self.optimizer.v), 0,
                                                  ac.num_params, C.addressof(hp), self.optimizer.step, L.ptr(self._norm_scratch),
                                                  L.ptr(self._grad_norm), L.stream_ptr()), "dwbc_clip_adam_step")
            if on_step is not None:
                on_step(k, "step")
            k += 1
        num_updates = self.num_learning_epochs * self.num_mini_batches
        losses = (self._losses / num_updates).tolist()                                   # single sync per update
        s.set_hist_latent(None)
        s.clear()
        value_mixing_ratio, priv_reg_coef = hp.mixing_ratio, hp.priv_reg_coef
        self.counter += 1                                                                 # PPO:259
        self.enforce_min_std()
        self.last_entropy = losses[3]
        ts_w = hp.torque_supervision_weight if self.torque_supervision else 0                            # PPO:158
        return losses[1], losses[0], (losses[4] if self.torque_supervision else 0.0), value_mixing_ratio, ts_w, losses[2], priv_reg_coef  # PPO:263

    def update_dagger(self, indices=None):
        """PPO:265-291."""
        ac, s, hp = self.actor_critic, self.storage, self._fill_hp()
        self._set_precision()
        self._packed = False
        if indices is None:
            indices, _ = s.draw_indices(self.num_mini_batches, self.generator)
        indices = indices.to(torch.int64).contiguous()
        mbs = indices.numel() // self.num_mini_batches
        ws = self._workspace(mbs)
        self._losses.zero_()
        hf, hc = ac.hist_range
        opt = self.hist_encoder_optimizer
        for batch_idx in s.mini_batch_generator(self.num_mini_batches, self.num_learning_epochs, indices):
            L.check(self._lib.dwbc_dagger_minibatch_grad(C.addressof(ac.net_cfg), L.ptr(ac.flat), s.c_struct_ptr(), L.ptr(batch_idx), mbs,
                                                         L.ptr(self.grad), L.ptr(self._losses), L.ptr(ws), L.stream_ptr()),
                    "dwbc_dagger_minibatch_grad")
            self._allreduce(hf, hc)
            opt.step += 1
            L.check(self._lib.dwbc_clip_adam_step(L.ptr(ac.flat), L.ptr(self.grad), L.ptr(opt.m), L.ptr(opt.v), hf, hc, C.addressof(hp),
                                                  opt.step, L.ptr(self._norm_scratch), None, L.stream_ptr()), "dwbc_clip_adam_step")
        num_updates = self.num_learning_epochs * self.num_mini_batches
        loss = float(self._losses[0]) / num_updates
        s.clear()
        self.counter += 1
        return loss

    def enforce_min_std(self):
        if self.min_policy_std is None:
            return
        ac = self.actor_critic
        self._packed = False
        L.check(self._lib.dwbc_enforce_min_std(L.ptr(ac.flat), ac.offsets["std"], L.ptr(self.min_policy_std),
                                               self.min_policy_std.numel(), L.stream_ptr()), "dwbc_enforce_min_std")
