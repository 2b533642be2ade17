"""Static description of the widowGo1 task consumed by the fused post-physics kernel.

Every default below restates a value of the reference configuration
(`legged_gym/legged_gym/envs/widowGo1/widowGo1_config.py`, cited per field as WGC:line;
base class `envs/base/legged_robot_config.py` as LRC:line) or a quantity the reference
derives at start-up from Isaac Gym (`envs/widowGo1/widowGo1.py` = WG:line).  When the
real legged_gym config object is available (`from_legged_gym`) the values are read from
it instead, so the kernel constants always follow the user's config.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

# Isaac Gym DOF order of the widowGo1 URDF (WG:1005 comment, WG:529/557 index conventions).
DOF_NAMES_IG = [
    "FL_hip_joint", "FL_thigh_joint", "FL_calf_joint",
    "FR_hip_joint", "FR_thigh_joint", "FR_calf_joint",
    "RL_hip_joint", "RL_thigh_joint", "RL_calf_joint",
    "RR_hip_joint", "RR_thigh_joint", "RR_calf_joint",
    "widow_waist", "widow_shoulder", "widow_elbow", "widow_forearm_roll",
    "widow_wrist_angle", "widow_wrist_rotate", "widow_left_finger", "widow_right_finger",
]
# "raisim"/hardware order used in observations and policy actions (WG:1015-1021).
DOF_NAMES_RAISIM = [
    "FR_hip_joint", "FR_thigh_joint", "FR_calf_joint",
    "FL_hip_joint", "FL_thigh_joint", "FL_calf_joint",
    "RR_hip_joint", "RR_thigh_joint", "RR_calf_joint",
    "RL_hip_joint", "RL_thigh_joint", "RL_calf_joint",
    "widow_waist", "widow_shoulder", "widow_elbow", "widow_forearm_roll",
    "widow_wrist_angle", "widow_wrist_rotate", "widow_left_finger", "widow_right_finger",
]

# ---------------------------------------------------------------------------------------
# Reward-term registry.  Ids are stable ABI (include/dwbc.h: DWBC_TERM_*).  The reference
# resolves terms by name through getattr(self, '_reward_' + name) (WG:141,158); the
# functions live at WG:1352-1469 and LR:832-922.  Terms whose reference implementation
# cannot run on widowGo1 (orientation: needs projected_gravity that WG:880-886 never
# computes; feet_stumble / arm_orientation: no such method) are deliberately absent.
# ---------------------------------------------------------------------------------------
REWARD_TERMS = [
    "action_rate", "ang_vel_xy", "arm_energy_abs_sum", "base_height", "collision", "dof_acc",
    "dof_pos_limits", "dof_vel", "dof_vel_limits", "energy_square", "feet_air_time",
    "feet_contact_forces", "foot_contacts_z", "hip_action_l2", "leg_action_l2", "leg_energy",
    "leg_energy_abs_sum", "leg_energy_sum_abs", "lin_vel_z", "stand_still", "stumble", "survive",
    "termination", "torque_limits", "torques", "tracking_ang_vel", "tracking_ang_vel_yaw_exp",
    "tracking_ang_vel_yaw_l1", "tracking_ee_cart", "tracking_ee_orn", "tracking_ee_orn_ry",
    "tracking_ee_sphere", "tracking_lin_vel", "tracking_lin_vel_x_exp", "tracking_lin_vel_x_l1",
    "tracking_lin_vel_y_l2", "tracking_lin_vel_z_l2",
]
TERM_ID = {n: i for i, n in enumerate(REWARD_TERMS)}
# episode_metric_sums keys, fixed order (WG:164).
METRIC_NAMES = ["leg_energy_abs_sum", "tracking_lin_vel_x_l1", "tracking_ang_vel_yaw_exp",
                "tracking_ee_cart", "tracking_ee_sphere", "tracking_ee_orn", "leg_action_l2",
                "torque", "energy_square", "foot_contacts_z"]
METRIC_ID = {n: i for i, n in enumerate(METRIC_NAMES)}
MAX_TERMS = 40          # active terms per channel the ABI struct can carry

# Column map of the dense per-step uniform table rand[N, RAND_COLS] (table mode) and of the
# Philox counter space (in-kernel mode).  One column per reference torch_rand_float element.
RAND_GOAL_ORN = 0        # 3   _resample_ee_goal_orn_once on timer expiry (WG:1307-1313)
RAND_GOAL_SPH = 3        # 30  up to 10 tries x (l,p,y)          (WG:1303-1306,1325-1330)
RAND_CMD = 33            # 2   _resample_commands in the callback (WG:837-839)
RAND_PUSH = 35           # 2   _push_robots                       (WG:808)
RAND_RST_DOF = 37        # 20  _reset_dofs                        (WG:824)
RAND_RST_XY = 57         # 2   _reset_root_states xy              (WG:767)
RAND_RST_VEL = 59        # 6   _reset_root_states velocities      (WG:774)
RAND_RST_CMD = 65        # 2   _resample_commands for timed-out envs (WG:726-727)
RAND_RST_GOAL_ORN = 67   # 3
RAND_RST_GOAL_SPH = 70   # 30
RAND_TERRAIN = 100       # 1   randint_like in _update_terrain_curriculum (LR:438)
RAND_COLS = 104


def _f32(x):
    return float(np.float32(x))


def _gains(table, dof_names, n_act):
    """WG:648-658: gain of the first key of cfg.control.stiffness / damping contained in the DOF name, 0 if none."""
    out = []
    for name in list(dof_names)[:n_act]:
        out.append(next((float(v) for k, v in table.items() if k in name), 0.0))
    return out


@dataclass
class WidowGo1Params:
    # ---- dimensions (WGC:117-125; n_bodies/gripper_idx are URDF-derived, WG:287,318) ----
    num_envs: int = 4096
    num_dofs: int = 20
    num_actions: int = 18
    num_bodies: int = 27
    gripper_idx: int = 24
    num_prop: int = 76
    num_priv: int = 24
    history_len: int = 10
    action_hist_len: int = 4            # action_delay + 2 (WG:541, WGC:120)
    action_delay: int = 2               # WGC:120; the policy action applied is action_history_buf[:, -action_delay - 1] (WG:1167-1168)
    clip_actions: float = 100.0         # WGC:116 normalization.clip_actions (WG:1163)
    feet_indices: List[int] = field(default_factory=lambda: [5, 9, 13, 17])   # FL,FR,RL,RR foot
    penalized_contact_indices: List[int] = field(default_factory=lambda: [1, 3, 7, 11, 15])
    termination_contact_indices: List[int] = field(default_factory=list)     # WGC:179 -> empty
    dof_names: List[str] = field(default_factory=lambda: list(DOF_NAMES_IG))
    reorder_dofs: bool = True           # WGC:129
    # ---- time (WG:80,118-119 with sim dt 0.005, decimation 4) ----
    dt: float = 0.02
    max_episode_length_s: float = 10.0  # WGC:127
    resampling_time: float = 3.0        # WGC:92
    push_interval_s: float = 3.0        # WGC:216
    push_robots: bool = True
    max_push_vel_xy: float = 0.5        # WGC:217
    # ---- normalisation (WGC:107-114) ----
    obs_scale_lin_vel: float = 1.0
    obs_scale_ang_vel: float = 1.0
    obs_scale_dof_pos: float = 1.0
    obs_scale_dof_vel: float = 0.05
    obs_scale_height: float = 5.0
    clip_observations: float = 100.0
    observe_priv: bool = True           # WGC:201 (precondition)
    # ---- init state / reset (WGC:131-160, LRC:85-89, WGC:313-314, WGC:186-192) ----
    default_dof_pos: List[float] = field(default_factory=lambda: [
        0.1, 0.8, -1.5, -0.1, 0.8, -1.5, 0.1, 0.8, -1.5, -0.1, 0.8, -1.5,
        0, 0, 0, 0, 0, 0, 0, 0])
    base_init_state: List[float] = field(default_factory=lambda: [
        0.0, 0.0, 0.42, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0])
    origin_perturb_range: float = 0.5
    init_vel_perturb_range: float = 0.1
    box_env_origins_x: float = 0.0
    box_env_origins_z: float = 0.1 / 2 + 0.16
    # ---- URDF-derived limits (LR:291-303); efforts follow SURVEY section 8d config 2 ----
    dof_pos_limits: List[List[float]] = field(default_factory=lambda:
        [[-0.863, 0.863], [-0.686, 4.501], [-2.818, -0.888]] * 4 +
        [[-3.14158, 3.14158], [-1.885, 1.972], [-2.147, 1.606], [-3.14158, 3.14158],
         [-1.745, 2.147], [-3.14158, 3.14158], [0.015, 0.037], [-0.037, -0.015]])
    dof_vel_limits: List[float] = field(default_factory=lambda: [30.1, 30.1, 20.06] * 4 +
                                        [3.14] * 6 + [1.0, 1.0])
    # PD controller of step() (WGC:166-170; p/d gains by DOF-name match WG:648-658): legs 50 / 1, arm 5 / 0.5
    p_gains: List[float] = field(default_factory=lambda: [50.0] * 12 + [5.0] * 6)
    d_gains: List[float] = field(default_factory=lambda: [1.0] * 12 + [0.5] * 6)
    action_scale: List[float] = field(default_factory=lambda: [0.4, 0.45, 0.45] * 4 + [2.1, 0.6, 0.6, 0.0, 0.0, 0.0])
    torque_limits: List[float] = field(default_factory=lambda: [23.7, 23.7, 35.55] * 4 +
                                       [10, 20, 15, 2, 5, 1, 0, 0])
    soft_dof_vel_limit: float = 1.0     # WGC:276
    soft_torque_limit: float = 1.0      # WGC:277
    # ---- termination (WG:945-948 hard-codes 0.2 rad; WGC:289 z threshold) ----
    term_roll: float = 0.2
    term_pitch: float = 0.2
    term_z: float = 0.325
    # ---- commands (WGC:90-105) ----
    lin_vel_x_clip: float = 0.3
    ang_vel_yaw_clip: float = 0.6
    init_lin_vel_x: List[float] = field(default_factory=lambda: [0.0, 0.0])
    final_lin_vel_x: List[float] = field(default_factory=lambda: [0.0, 0.9])
    init_ang_vel_yaw: List[float] = field(default_factory=lambda: [0.0, 0.0])
    final_ang_vel_yaw: List[float] = field(default_factory=lambda: [-1.0, 1.0])
    final_tracking_ang_vel_yaw_exp: float = 0.15
    lin_vel_x_schedule: List[float] = field(default_factory=lambda: [0, 1])
    ang_vel_yaw_schedule: List[float] = field(default_factory=lambda: [0, 1])
    tracking_ang_vel_yaw_schedule: List[float] = field(default_factory=lambda: [0, 1])
    # ---- EE goal generator (WGC:46-82) ----
    traj_time: List[float] = field(default_factory=lambda: [1.0, 3.0])
    hold_time: List[float] = field(default_factory=lambda: [0.5, 2.0])
    collision_upper_limits: List[float] = field(default_factory=lambda: [0.3, 0.15, 0.05 - 0.165])
    collision_lower_limits: List[float] = field(default_factory=lambda: [-0.2, -0.15, -0.35 - 0.165])
    underground_limit: float = -0.57
    num_collision_check_samples: int = 10
    command_mode: str = "sphere"
    init_pos_l: List[float] = field(default_factory=lambda: [0.6, 0.6])
    init_pos_p: List[float] = field(default_factory=lambda: [math.pi / 4, math.pi / 4])
    init_pos_y: List[float] = field(default_factory=lambda: [-math.pi / 6, math.pi / 6])
    final_pos_l: List[float] = field(default_factory=lambda: [0.2, 0.7])
    final_pos_p: List[float] = field(default_factory=lambda: [-2 * math.pi / 5, 1 * math.pi / 5])
    final_pos_y: List[float] = field(default_factory=lambda: [-3 * math.pi / 5, 3 * math.pi / 5])
    final_delta_orn: List[List[float]] = field(default_factory=lambda: [[0, 0], [0, 0], [0, 0]])
    final_tracking_ee_reward: float = 0.55
    l_schedule: List[float] = field(default_factory=lambda: [0, 1])
    p_schedule: List[float] = field(default_factory=lambda: [0, 1])
    y_schedule: List[float] = field(default_factory=lambda: [0, 1])
    tracking_ee_reward_schedule: List[float] = field(default_factory=lambda: [0, 1])
    orn_error_scale: List[float] = field(default_factory=lambda: [2 / math.pi] * 3)
    z_invariant_offset: float = 0.53    # WG:597
    # ---- rewards (WGC:226-279) ----
    reward_scales: Dict[str, float] = field(default_factory=lambda: {
        "energy_square": -6e-5, "foot_contacts_z": -1e-4, "hip_action_l2": -0.01,
        "survive": 0.2, "tracking_ang_vel_yaw_exp": 0.15, "tracking_lin_vel_x_l1": 0.5})
    arm_reward_scales: Dict[str, float] = field(default_factory=lambda: {
        "arm_energy_abs_sum": -0.004, "tracking_ee_sphere": 0.55})
    only_positive_rewards: bool = False
    tracking_sigma: float = 1.0
    tracking_ee_sigma: float = 1.0
    base_height_target: float = 0.25
    max_contact_force: float = 100.0
    # ---- terrain / height scan (WGC:291-314, LR:777-829) ----
    measure_heights: bool = False
    measured_points_x: List[float] = field(default_factory=lambda: [
        -0.8, -0.7, -0.6, -0.5, -0.4, -0.3, -0.2, -0.1, 0., 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8])
    measured_points_y: List[float] = field(default_factory=lambda: [
        -0.5, -0.4, -0.3, -0.2, -0.1, 0., 0.1, 0.2, 0.3, 0.4, 0.5])
    horizontal_scale: float = 0.025
    vertical_scale: float = 1 / 100000
    border_size: float = 0.0
    tot_rows: int = 10000
    tot_cols: int = 600
    terrain_curriculum: bool = False
    terrain_env_length: float = 8.0     # LRC:56 terrain_length
    max_terrain_level: int = 10         # LRC:59 num_rows
    terrain_num_cols: int = 20          # LRC:60

    # ------------------------------------------------------------------------------------
    @property
    def num_obs(self) -> int:
        return self.num_prop * (self.history_len + 1) + self.num_priv

    @property
    def max_episode_length(self) -> float:
        return float(np.ceil(self.max_episode_length_s / self.dt))      # WG:118

    @property
    def push_interval(self) -> float:
        return float(np.ceil(self.push_interval_s / self.dt))           # WG:119

    @property
    def resample_interval(self) -> int:
        return int(self.resampling_time / self.dt)                      # WG:922

    @property
    def sphere_error_scale(self) -> List[float]:
        return [1 / (self.final_pos_l[1] - self.final_pos_l[0]),
                1 / (self.final_pos_p[1] - self.final_pos_p[0]),
                1 / (self.final_pos_y[1] - self.final_pos_y[0])]         # WGC:76

    @property
    def num_height_points(self) -> int:
        return len(self.measured_points_x) * len(self.measured_points_y)

    def ig2raisim(self, n=None) -> List[int]:
        """obs column j <- Isaac Gym dof index (WG:1010-1028)."""
        names = DOF_NAMES_RAISIM[: (n or self.num_dofs)]
        if not self.reorder_dofs:
            return list(range(len(names)))
        return [self.dof_names.index(nm) for nm in names]

    def raisim2ig(self, n=None) -> List[int]:
        """Isaac Gym dof i <- policy action index (WG:1030-1048, 1070-1088)."""
        n = n or self.num_dofs
        if not self.reorder_dofs:
            return list(range(n))
        return [DOF_NAMES_RAISIM[:n].index(nm) for nm in self.dof_names[:n]]

    def feet_perm(self) -> List[int]:
        return [1, 0, 3, 2] if self.reorder_dofs else [0, 1, 2, 3]     # WG:1003-1008

    def active_terms(self, which: str) -> List[str]:
        """Alphabetical list of non-zero, non-'termination' terms (WG:123-160).

        class_to_dict iterates dir(obj) (`utils/helpers.py:44-56`), i.e. sorted names; the
        order fixes the fp32 summation order of rew_buf (WG:176-181)."""
        d = self.reward_scales if which == "leg" else self.arm_reward_scales
        return [k for k in sorted(d) if d[k] != 0 and k != "termination"]

    def sum_slots(self) -> List[str]:
        """Keys of episode_sums in reference order (WG:160-161): leg dict keys then arm dict
        keys, each including a non-zero 'termination'."""
        leg = [k for k in sorted(self.reward_scales) if self.reward_scales[k] != 0]
        arm = [k for k in sorted(self.arm_reward_scales) if self.arm_reward_scales[k] != 0]
        out = []
        for k in leg + arm:
            if k not in out:
                out.append(k)
        return out

    # curriculum (WG:675-692): host-side linear schedules.
    @staticmethod
    def curriculum_value(schedule, init, final, counter):
        init = np.asarray(init, dtype=np.float64)
        final = np.asarray(final, dtype=np.float64)
        return np.clip((counter - schedule[0]) / (schedule[1] - schedule[0]), 0, 1) * (final - init) + init

    @classmethod
    def from_legged_gym(cls, cfg, *, num_envs, dt, dof_names, num_bodies, gripper_idx, feet_indices,
                        penalized_contact_indices, termination_contact_indices, dof_pos_limits,
                        dof_vel_limits, torque_limits, default_dof_pos, base_init_state,
                        reward_scales, arm_reward_scales):
        """Build from a live legged_gym `WidowGo1RoughCfg` plus the URDF-derived tensors the
        reference computes in `_create_envs` / `_process_dof_props` / `_init_buffers`."""
        g, c, t = cfg.goal_ee, cfg.commands, cfg.terrain
        p = cls(
            num_envs=num_envs, num_dofs=len(dof_names), num_actions=cfg.env.num_actions,
            num_bodies=num_bodies, gripper_idx=gripper_idx, num_prop=cfg.env.num_proprio,
            num_priv=cfg.env.num_priv, history_len=cfg.env.history_len,
            action_hist_len=cfg.env.action_delay + 2, action_delay=cfg.env.action_delay, clip_actions=cfg.normalization.clip_actions,
            feet_indices=list(feet_indices),
            penalized_contact_indices=list(penalized_contact_indices),
            termination_contact_indices=list(termination_contact_indices),
            dof_names=list(dof_names), reorder_dofs=cfg.env.reorder_dofs, dt=dt,
            max_episode_length_s=cfg.env.episode_length_s, resampling_time=c.resampling_time,
            push_interval_s=cfg.domain_rand.push_interval_s, push_robots=cfg.domain_rand.push_robots,
            max_push_vel_xy=cfg.domain_rand.max_push_vel_xy,
            obs_scale_lin_vel=cfg.normalization.obs_scales.lin_vel,
            obs_scale_ang_vel=cfg.normalization.obs_scales.ang_vel,
            obs_scale_dof_pos=cfg.normalization.obs_scales.dof_pos,
            obs_scale_dof_vel=cfg.normalization.obs_scales.dof_vel,
            obs_scale_height=cfg.normalization.obs_scales.height_measurements,
            clip_observations=cfg.normalization.clip_observations,
            observe_priv=cfg.domain_rand.observe_priv,
            default_dof_pos=[float(x) for x in default_dof_pos],
            p_gains=_gains(cfg.control.stiffness, dof_names, cfg.env.num_actions), d_gains=_gains(cfg.control.damping, dof_names, cfg.env.num_actions),
            action_scale=[float(x) for x in cfg.control.action_scale],
            base_init_state=[float(x) for x in base_init_state],
            origin_perturb_range=t.origin_perturb_range, init_vel_perturb_range=t.init_vel_perturb_range,
            box_env_origins_x=cfg.box.box_env_origins_x, box_env_origins_z=cfg.box.box_env_origins_z,
            dof_pos_limits=[[float(a), float(b)] for a, b in dof_pos_limits],
            dof_vel_limits=[float(x) for x in dof_vel_limits],
            torque_limits=[float(x) for x in torque_limits],
            soft_dof_vel_limit=cfg.rewards.soft_dof_vel_limit, soft_torque_limit=cfg.rewards.soft_torque_limit,
            term_z=cfg.termination.z_threshold,
            lin_vel_x_clip=c.lin_vel_x_clip, ang_vel_yaw_clip=c.ang_vel_yaw_clip,
            init_lin_vel_x=list(c.ranges.init_lin_vel_x), final_lin_vel_x=list(c.ranges.final_lin_vel_x),
            init_ang_vel_yaw=list(c.ranges.init_ang_vel_yaw), final_ang_vel_yaw=list(c.ranges.final_ang_vel_yaw),
            final_tracking_ang_vel_yaw_exp=c.ranges.final_tracking_ang_vel_yaw_exp,
            lin_vel_x_schedule=list(c.lin_vel_x_schedule), ang_vel_yaw_schedule=list(c.ang_vel_yaw_schedule),
            tracking_ang_vel_yaw_schedule=list(c.tracking_ang_vel_yaw_schedule),
            traj_time=list(g.traj_time), hold_time=list(g.hold_time),
            collision_upper_limits=list(g.collision_upper_limits),
            collision_lower_limits=list(g.collision_lower_limits),
            underground_limit=g.underground_limit, num_collision_check_samples=g.num_collision_check_samples,
            command_mode=g.command_mode,
            init_pos_l=list(g.ranges.init_pos_l), init_pos_p=list(g.ranges.init_pos_p),
            init_pos_y=list(g.ranges.init_pos_y), final_pos_l=list(g.ranges.final_pos_l),
            final_pos_p=list(g.ranges.final_pos_p), final_pos_y=list(g.ranges.final_pos_y),
            final_delta_orn=[list(r) for r in g.ranges.final_delta_orn],
            final_tracking_ee_reward=g.ranges.final_tracking_ee_reward,
            l_schedule=list(g.l_schedule), p_schedule=list(g.p_schedule), y_schedule=list(g.y_schedule),
            tracking_ee_reward_schedule=list(g.tracking_ee_reward_schedule),
            orn_error_scale=list(g.orn_error_scale),
            reward_scales=dict(reward_scales), arm_reward_scales=dict(arm_reward_scales),
            only_positive_rewards=cfg.rewards.only_positive_rewards,
            tracking_sigma=cfg.rewards.tracking_sigma, tracking_ee_sigma=cfg.rewards.tracking_ee_sigma,
            base_height_target=cfg.rewards.base_height_target, max_contact_force=cfg.rewards.max_contact_force,
            measure_heights=t.measure_heights, measured_points_x=list(t.measured_points_x),
            measured_points_y=list(t.measured_points_y), horizontal_scale=t.horizontal_scale,
            vertical_scale=t.vertical_scale, border_size=t.border_size,
            tot_rows=getattr(t, "tot_rows", 0), tot_cols=getattr(t, "tot_cols", 0),
            terrain_curriculum=t.curriculum,
        )
        return p


class CommandCurriculum:
    """Host-side schedules of `WidowGo1.update_command_curriculum` (WG:678-692).

    Produces the runtime ranges / reward scales that are passed to the kernel per step
    (they are kernel *arguments*, not constants: SURVEY section 5 'config / flags')."""

    def __init__(self, p: WidowGo1Params):
        self.p = p
        self.update_counter = 0
        self.lin_vel_x_ranges = np.array(p.init_lin_vel_x, dtype=np.float64)
        self.ang_vel_yaw_ranges = np.array(p.init_ang_vel_yaw, dtype=np.float64)
        self.goal_ee_l_ranges = np.array(p.init_pos_l, dtype=np.float64)
        self.goal_ee_p_ranges = np.array(p.init_pos_p, dtype=np.float64)
        self.goal_ee_y_ranges = np.array(p.init_pos_y, dtype=np.float64)
        self.reward_scales = dict(p.reward_scales)
        self.arm_reward_scales = dict(p.arm_reward_scales)

    def update(self):
        p, cv = self.p, WidowGo1Params.curriculum_value
        self.update_counter += 1
        n = self.update_counter
        self.lin_vel_x_ranges = cv(p.lin_vel_x_schedule, p.init_lin_vel_x, p.final_lin_vel_x, n)
        self.ang_vel_yaw_ranges = cv(p.ang_vel_yaw_schedule, p.init_ang_vel_yaw, p.final_ang_vel_yaw, n)
        self.reward_scales["tracking_ang_vel_yaw_exp"] = float(
            cv(p.tracking_ang_vel_yaw_schedule, 0, p.final_tracking_ang_vel_yaw_exp, n))
        self.goal_ee_l_ranges = cv(p.l_schedule, p.init_pos_l, p.final_pos_l, n)
        self.goal_ee_p_ranges = cv(p.p_schedule, p.init_pos_p, p.final_pos_p, n)
        self.goal_ee_y_ranges = cv(p.y_schedule, p.init_pos_y, p.final_pos_y, n)
        key = "tracking_ee_sphere" if "tracking_ee_sphere" in self.arm_reward_scales else "tracking_ee_cart"
        self.arm_reward_scales[key] = float(
            cv(p.tracking_ee_reward_schedule, 0, p.final_tracking_ee_reward, n))
