"""Shared test helpers: build oracle / kernel env state from the synthetic factories."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import dwbc_b200  # noqa: E402,F401
from dwbc_b200 import synth  # noqa: E402

ENV_CONFIGS = {
    # default widowGo1 flat config (BASELINE.json configs[1] semantics)
    "flat": dict(),
    # BASELINE.json configs[2]: flat reward set + 187-point height scan on the widowGo1 field (WG:253: 10000 x 600 int16) + terrain curriculum
    "rough": dict(measure_heights=True, tot_rows=10000, tot_cols=600, terrain_curriculum=True),
    # every optional branch: extra reward terms on both channels, termination terms, contact
    # termination, positive-reward clip off, height scan on, terrain curriculum on (LR:421-441)
    "full": dict(
        measure_heights=True, tot_rows=400, tot_cols=600, termination_contact_indices=[2], terrain_curriculum=True,
        reward_scales={
            "action_rate": -0.01, "ang_vel_xy": -0.05, "base_height": -1.0, "collision": -1.0, "dof_acc": -2.5e-7,
            "dof_pos_limits": -10.0, "dof_vel": -1e-3, "dof_vel_limits": -0.1, "energy_square": -6e-5,
            "feet_air_time": 1.0, "feet_contact_forces": -0.01, "foot_contacts_z": -1e-4, "hip_action_l2": -0.01,
            "leg_action_l2": -0.005, "leg_energy": -1e-3, "leg_energy_abs_sum": -1e-3, "leg_energy_sum_abs": -1e-3,
            "lin_vel_z": -2.0, "stand_still": -0.1, "stumble": -0.5, "survive": 0.2, "termination": -2.0,
            "torque_limits": -0.01, "torques": -1e-5, "tracking_ang_vel": 0.5, "tracking_ang_vel_yaw_exp": 0.15,
            "tracking_ang_vel_yaw_l1": 0.1, "tracking_lin_vel": 1.0, "tracking_lin_vel_x_exp": 0.2,
            "tracking_lin_vel_x_l1": 0.5, "tracking_lin_vel_y_l2": -0.1, "tracking_lin_vel_z_l2": -0.1},
        arm_reward_scales={
            "arm_energy_abs_sum": -0.004, "termination": -1.0, "tracking_ee_cart": 0.3, "tracking_ee_orn": 0.1,
            "tracking_ee_orn_ry": 0.1, "tracking_ee_sphere": 0.55}),
}


def make_params(name, num_envs):
    return dwbc_b200.WidowGo1Params(num_envs=num_envs, **ENV_CONFIGS[name])


def runtime(p, counter=1):
    """Curriculum outputs after `counter` calls of update_command_curriculum (WG:678-692)."""
    cur = dwbc_b200.CommandCurriculum(p)
    for _ in range(counter):
        cur.update()
    return SimpleNamespace(lin_vel_x=cur.lin_vel_x_ranges, ang_vel_yaw=cur.ang_vel_yaw_ranges,
                           goal_l=cur.goal_ee_l_ranges, goal_p=cur.goal_ee_p_ranges, goal_y=cur.goal_ee_y_ranges,
                           leg_scales=cur.reward_scales, arm_scales=cur.arm_reward_scales)


def initial(p, seed):
    st = synth.initial_env_state(p, seed)
    st.update(synth.sim_state(p, seed, 0))
    if p.measure_heights:
        st["height_samples"] = synth.height_field(p, seed)
    return st


def sim_state(p, seed, t, env_origins=None, **kw):
    """synth.sim_state, with the robot placed RELATIVE to its current env origin when the terrain curriculum is on
    (LR:430 measures the distance walked from the origin: absolute +-5 m positions would make every reset a promotion).
    `env_origins` = the [N,3] origins before the step (numpy / tensor), i.e. state the caller carries."""
    sim = synth.sim_state(p, seed, t, **kw)
    if p.terrain_curriculum and env_origins is not None:
        org = env_origins.detach().cpu().numpy() if isinstance(env_origins, torch.Tensor) else np.asarray(env_origins)
        sim["root_states"][:, 0, 0:2] += org[:, 0:2].astype(np.float32)
    return sim


def oracle_state(p, st):
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).clone()  # noqa: E731
    N = p.num_envs
    s = SimpleNamespace(
        root_states_full=T(st["root_states"]), dof_state=T(st["dof_state"]), rigid_body_state=T(st["rigid_body_state"]),
        contact_forces_full=T(st["contact_forces"]), force_sensor=T(st["force_sensor"]), torques=T(st["torques"]),
        action_history_buf=T(st["action_history_buf"]),
        base_lin_vel=torch.zeros(N, 3), base_ang_vel=torch.zeros(N, 3), base_yaw_euler=torch.zeros(N, 3),
        base_yaw_quat=torch.zeros(N, 4), mass_params=T(st["mass_params"]), friction=T(st["friction"]),
        motor_strength=T(st["motor_strength"]))
    for k in ("commands", "goal_timer", "traj_timesteps", "traj_total_timesteps", "ee_start_sphere", "ee_goal_sphere",
              "ee_goal_cart", "curr_ee_goal_sphere", "curr_ee_goal_cart", "ee_goal_delta_orn_euler",
              "ee_goal_orn_euler", "obs_history_buf", "last_actions", "last_dof_vel", "last_root_vel", "feet_air_time",
              "last_contacts", "env_origins", "box_env_origins_delta_y", "episode_length_buf", "terrain_levels",
              "terrain_types", "terrain_origins"):
        setattr(s, k, T(st[k]))
    s.actions = s.action_history_buf[:, -3].clone()
    if "height_samples" in st:
        s.height_samples = T(st["height_samples"])
    return s


def load_sim_into_oracle(o, p, sim):
    s = o.s
    s.root_states_full.copy_(torch.from_numpy(sim["root_states"]))
    s.dof_state.copy_(torch.from_numpy(sim["dof_state"]))
    s.rigid_body_state.copy_(torch.from_numpy(sim["rigid_body_state"]))
    s.contact_forces_full.copy_(torch.from_numpy(sim["contact_forces"]))
    s.force_sensor.copy_(torch.from_numpy(sim["force_sensor"]))
    s.torques = torch.from_numpy(sim["torques"]).clone()
    a = torch.from_numpy(sim["policy_actions"])[:, p.raisim2ig(p.num_actions)]
    a = torch.clip(a, -100.0, 100.0)
    s.action_history_buf = torch.cat([s.action_history_buf[:, 1:], a[:, None, :]], dim=1)
    s.actions = s.action_history_buf[:, -3].clone()
