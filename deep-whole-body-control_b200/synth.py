"""Deterministic synthetic inputs (sim-state tensors, rollout storage, network parameters).

Isaac Gym cannot run here, and the metric is defined on synthetic
sim-state tensors (BASELINE.json north_star).  Everything is
derived from a counter-based integer hash (splitmix64 finaliser) evaluated with numpy
uint64 arithmetic, so the same arrays are produced on every machine without depending on
torch / numpy RNG stream stability.  Golden fixtures (tests/golden) store only outputs;
their inputs are regenerated from (seed, stream) by these functions.
"""
from __future__ import annotations

import math

import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)


def _mix(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    x ^= x >> np.uint64(30)
    x *= _M1
    x ^= x >> np.uint64(27)
    x *= _M2
    x ^= x >> np.uint64(31)
    return x


def uniform(seed: int, stream: int, shape, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    """float32 uniforms in [lo, hi): 24-bit mantissa values k * 2^-24 (exact)."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        base = _mix(np.array([seed], dtype=np.uint64) * _G + np.array([stream], dtype=np.uint64) * _M1)
        idx = np.arange(n, dtype=np.uint64) * _G + base
        bits = _mix(idx) >> np.uint64(40)
    u = bits.astype(np.float32) * np.float32(2.0 ** -24)
    if lo != 0.0 or hi != 1.0:
        u = (np.float32(hi - lo) * u + np.float32(lo)).astype(np.float32)
    return u.reshape(shape)


def normal(seed: int, stream: int, shape, mean: float = 0.0, std: float = 1.0) -> np.ndarray:
    """Approximately N(mean, std): Irwin-Hall sum of 4 uniforms (exact fp32 arithmetic)."""
    shape = tuple(shape) if hasattr(shape, "__len__") else (int(shape),)
    u = uniform(seed, stream, (4,) + shape)
    z = ((u[0] + u[1]) + (u[2] + u[3]) - np.float32(2.0)) * np.float32(math.sqrt(3.0))
    return (z * np.float32(std) + np.float32(mean)).astype(np.float32)


def bernoulli(seed: int, stream: int, shape, p: float) -> np.ndarray:
    return uniform(seed, stream, shape) < np.float32(p)


# ---------------------------------------------------------------------------------------
# sim-state factory
# ---------------------------------------------------------------------------------------

def _quat_from_rpy(r, p, y):
    cy, sy = np.cos(y * 0.5), np.sin(y * 0.5)
    cr, sr = np.cos(r * 0.5), np.sin(r * 0.5)
    cp, sp = np.cos(p * 0.5), np.sin(p * 0.5)
    q = np.stack([cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp,
                  sy * cr * cp - cy * sr * sp, cy * cr * cp + sy * sr * sp], axis=-1)
    return q.astype(np.float32)


def sim_state(p, seed: int, step: int, rp_sigma: float = 0.1, z_lo: float = 0.30) -> dict:
    """One step's worth of Isaac-Gym-layout tensors for N envs (numpy float32).

    Layouts follow WG:505-558: `_root_states[N,2,13]` (robot, box), `dof_state[N*n_dof,2]`,
    `_rigid_body_state[N,n_body+1,13]`, `_contact_forces[N,n_body+1,3]`,
    `force_sensor_tensor[N,4,6]`, `torques[N,n_dof]`.  `policy_actions[N,18]` are raw policy
    outputs in raisim order (input of `step`, WG:1162)."""
    N, nd, nb = p.num_envs, p.num_dofs, p.num_bodies
    s = 1000 * step
    root = np.zeros((N, 2, 13), np.float32)
    root[:, 0, 0:2] = uniform(seed, s + 1, (N, 2), -5, 5)
    root[:, 0, 2] = uniform(seed, s + 2, (N,), z_lo, 0.50)
    rpy = normal(seed, s + 3, (N, 2), 0.0, rp_sigma)
    yaw = uniform(seed, s + 4, (N,), -math.pi, math.pi)
    root[:, 0, 3:7] = _quat_from_rpy(rpy[:, 0], rpy[:, 1], yaw)
    root[:, 0, 7:13] = normal(seed, s + 5, (N, 6), 0.0, 0.5)
    root[:, 1, 0:3] = uniform(seed, s + 6, (N, 3), -1, 1)
    root[:, 1, 6] = 1.0
    dof = np.zeros((N, nd, 2), np.float32)
    dof[:, :, 0] = np.asarray(p.default_dof_pos, np.float32) + normal(seed, s + 7, (N, nd), 0.0, 0.3)
    dof[:, 12, 0] += uniform(seed, s + 8, (N,), -4, 4)          # exercise the waist wrap (WG:970)
    dof[:, :, 1] = normal(seed, s + 9, (N, nd), 0.0, 2.0)
    rb = normal(seed, s + 10, (N, nb + 1, 13), 0.0, 1.0)
    cy, sy = np.cos(yaw), np.sin(yaw)
    loc = uniform(seed, s + 11, (N, 3), 0.2, 0.7)
    loc[:, 1] -= 0.45
    g = p.gripper_idx
    rb[:, g, 0] = root[:, 0, 0] + cy * loc[:, 0] - sy * loc[:, 1]
    rb[:, g, 1] = root[:, 0, 1] + sy * loc[:, 0] + cy * loc[:, 1]
    rb[:, g, 2] = 0.53 + loc[:, 2] - 0.45
    q = rb[:, :, 3:7]
    rb[:, :, 3:7] = q / np.linalg.norm(q, axis=-1, keepdims=True)
    cf = normal(seed, s + 12, (N, nb + 1, 3), 0.0, 0.3)
    cf[:, :, 2] = np.abs(cf[:, :, 2]) * 2.0
    fs = normal(seed, s + 13, (N, 4, 6), 0.0, 0.9)
    lim = np.asarray(p.torque_limits, np.float32)
    tq = np.clip(normal(seed, s + 14, (N, nd), 0.0, 5.0), -lim, lim)
    act = normal(seed, s + 15, (N, p.num_actions), 0.0, 1.0)
    return dict(root_states=root, dof_state=dof.reshape(N * nd, 2), rigid_body_state=rb.astype(np.float32),
                contact_forces=cf.astype(np.float32), force_sensor=fs, torques=tq.astype(np.float32),
                policy_actions=act)


def env_static(p, seed: int) -> dict:
    """Per-env constants the reference draws once at start-up (WG:455, 476-484, 218-227, 574-575)."""
    N = p.num_envs
    mass = np.concatenate([uniform(seed, 101, (N, 1), -0.5, 2.5), uniform(seed, 102, (N, 3), -0.15, 0.15),
                           uniform(seed, 103, (N, 1), 0.0, 0.1)], axis=1)
    traj = uniform(seed, 107, (N,), p.traj_time[0], p.traj_time[1]) / np.float32(p.dt)
    hold = uniform(seed, 108, (N,), p.hold_time[0], p.hold_time[1]) / np.float32(p.dt)
    org = np.zeros((N, 3), np.float32)
    org[:, 0] = uniform(seed, 109, (N,), -3.75, -3.0)
    org[:, 1] = uniform(seed, 110, (N,), -100, 100)
    sign = np.where(bernoulli(seed, 111, (N,), 0.5), 1.0, -1.0).astype(np.float32)
    return dict(mass_params=mass.astype(np.float32), friction=uniform(seed, 104, (N, 1), -0.5, 3.0),
                motor_strength=uniform(seed, 105, (N, p.num_actions), 0.7, 1.3),
                traj_timesteps=traj.astype(np.float32), traj_total_timesteps=(traj + hold).astype(np.float32),
                env_origins=org, box_env_origins_delta_y=sign * uniform(seed, 112, (N,), 0.1, 0.3))


def height_field(p, seed: int) -> np.ndarray:
    """int16 `height_samples[tot_rows, tot_cols]` (LR:793-829 input; WG:253 layout): smooth bumps
    quantised by vertical_scale, clipped to the int16 range."""
    rows, cols = p.tot_rows, p.tot_cols
    x = np.arange(rows, dtype=np.float32)[:, None] * np.float32(p.horizontal_scale)
    y = np.arange(cols, dtype=np.float32)[None, :] * np.float32(p.horizontal_scale)
    z = 0.08 * np.sin(1.7 * x) * np.cos(2.3 * y) + 0.04 * np.sin(5.1 * x + 0.7 * y)
    z = z + 0.01 * (uniform(seed, 201, (rows, cols)) - 0.5)
    return np.clip(np.round(z / p.vertical_scale), -32768, 32767).astype(np.int16)


# ---------------------------------------------------------------------------------------
# PPO inputs
# ---------------------------------------------------------------------------------------

def rollout_inputs(N: int, T: int, n_obs: int, seed: int) -> dict:
    """obs~N(0,1), rewards~N(0,1), dones~Bern(0.05), time_outs~Bern(0.02)."""
    return dict(obs=normal(seed, 301, (T + 1, N, n_obs)), rew=normal(seed, 302, (T, N)),
                arm_rew=normal(seed, 303, (T, N)), dones=bernoulli(seed, 304, (T, N), 0.05),
                time_outs=bernoulli(seed, 305, (T, N), 0.02), eps=normal(seed, 306, (T, N, 18)))


def arm_torque_inputs(N: int, T: int, n_arm: int, seed: int) -> dict:
    """Targets of the arm torque-supervision branch (PPO:136-142, RS:82-84): operational-space torques ~ N(0, 3), arm joint
    positions ~ U(-1.5, 1.5), velocities ~ N(0, 1); `coefs` = what OPR:91 hands to set_arm_default_coeffs (widow gains WGC:166-167
    with a per-joint spread so that a swapped coefficient shows, default joint positions ~ U(-0.5, 0.5))."""
    j = np.arange(n_arm, dtype=np.float32)
    return dict(target_arm_torques=normal(seed, 311, (T, N, n_arm), 0.0, 3.0), current_arm_dof_pos=uniform(seed, 312, (T, N, n_arm), -1.5, 1.5),
                current_arm_dof_vel=normal(seed, 313, (T, N, n_arm)),
                coefs=(np.float32(5.0) + np.float32(0.25) * j, np.float32(0.5) + np.float32(0.05) * j, uniform(seed, 314, (n_arm,), -0.5, 0.5)))


def policy_params(shapes, seed: int) -> list:
    """U(-1/sqrt(fan_in), 1/sqrt(fan_in)) per tensor (the bound nn.Linear/Conv1d default init
    uses, `rsl_rl/modules/actor_critic.py` relies on torch defaults); `shapes` = [(name, shape)]."""
    out = []
    for i, (name, shape) in enumerate(shapes):
        if name == "std":
            out.append(None)
            continue
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else None
        if fan_in is None:            # bias: bound from the matching weight, which precedes it
            fan_in = int(np.prod(shapes[i - 1][1][1:]))
        b = 1.0 / math.sqrt(fan_in)
        out.append(uniform(seed, 400 + i, shape, -b, b))
    return out


def initial_env_state(p, seed: int, random_ep_len: bool = True) -> dict:
    """Task-owned state at t=0 in the reference's layouts/dtypes (WG:498-672, BT:71-83).

    Mid-training-like values (non-trivial goals, history, timers, episode lengths as after
    `init_at_random_ep_len`, OPR:107-108) so a short trajectory exercises every branch."""
    N, H, P = p.num_envs, p.history_len, p.num_prop
    st = env_static(p, seed)
    lo = [p.final_pos_l[0], p.final_pos_p[0], p.final_pos_y[0]]
    hi = [p.final_pos_l[1], p.final_pos_p[1], p.final_pos_y[1]]
    start = np.stack([uniform(seed, 120 + i, (N,), lo[i], hi[i]) for i in range(3)], axis=-1)
    goal = np.stack([uniform(seed, 123 + i, (N,), lo[i], hi[i]) for i in range(3)], axis=-1)
    ep = (uniform(seed, 130, (N,)) * np.float32(p.max_episode_length + 2)).astype(np.int64) if random_ep_len \
        else np.zeros((N,), np.int64)
    if random_ep_len:                       # a few envs about to time out / resample commands
        k = min(N, 8)
        ep[:k] = np.array([499, 500, 498, 497, 149, 148, 299, 0], np.int64)[:k]
    cmd = np.zeros((N, 3), np.float32)
    cmd[:, 0] = uniform(seed, 131, (N,), 0.0, 0.9) * bernoulli(seed, 132, (N,), 0.7)
    cmd[:, 2] = uniform(seed, 133, (N,), -1.0, 1.0) * bernoulli(seed, 134, (N,), 0.7)
    timer = np.floor(uniform(seed, 135, (N,)) * (st["traj_total_timesteps"] + 2)).astype(np.float32)
    st.update(dict(
        commands=cmd, goal_timer=timer,
        ee_start_sphere=start.astype(np.float32), ee_goal_sphere=goal.astype(np.float32),
        ee_goal_cart=np.zeros((N, 3), np.float32), curr_ee_goal_sphere=start.astype(np.float32).copy(),
        curr_ee_goal_cart=np.zeros((N, 3), np.float32),
        ee_goal_delta_orn_euler=np.zeros((N, 3), np.float32), ee_goal_orn_euler=np.zeros((N, 3), np.float32),
        obs_history_buf=normal(seed, 140, (N, H, P), 0.0, 0.5),
        action_history_buf=normal(seed, 141, (N, p.action_hist_len, p.num_actions), 0.0, 1.0),
        episode_length_buf=ep,
        last_actions=normal(seed, 142, (N, p.num_actions)), last_dof_vel=normal(seed, 143, (N, p.num_dofs)),
        last_root_vel=np.zeros((N, 6), np.float32), feet_air_time=uniform(seed, 144, (N, 4), 0.0, 0.4),
        last_contacts=bernoulli(seed, 145, (N, 4), 0.5),
        terrain_levels=(uniform(seed, 146, (N,)) * p.max_terrain_level).astype(np.int64),
        terrain_types=(np.arange(N) * p.terrain_num_cols // max(N, 1)).astype(np.int64),
    ))
    tl, tc = p.max_terrain_level, p.terrain_num_cols
    org = np.zeros((tl, tc, 3), np.float32)
    org[:, :, 0] = (np.arange(tl, dtype=np.float32)[:, None] + 0.5) * np.float32(p.terrain_env_length)
    org[:, :, 1] = (np.arange(tc, dtype=np.float32)[None, :] + 0.5) * np.float32(p.terrain_env_length)
    st["terrain_origins"] = org
    return st


def rand_table(p, seed: int, step: int) -> np.ndarray:
    from .config import RAND_COLS
    return uniform(seed, 5000 + step, (p.num_envs, RAND_COLS))
