"""TEST INFRASTRUCTURE ONLY -- CPU (torch fp32) restatement of the widowGo1 post-physics path.

Follows, in order, `WidowGo1.post_physics_step` (legged_gym/legged_gym/envs/widowGo1/
widowGo1.py:865-915, cited WG:line) and the functions it calls; `LeggedRobot._get_heights`
and `_update_terrain_curriculum` (envs/base/legged_robot.py, cited LR:line).  It is a
restatement in a functional style over a plain namespace of tensors, not a copy: the
reference's sparse `env_ids` scatter chains are expressed as dense masked updates driven by
one pre-drawn uniform table ``rand[N, RAND_COLS]`` (column map in
deep-whole-body-control_b200/config.py) so that the CUDA kernel, this oracle and the
reference's own Python (executed through tests/fakes with torch_rand_float redirected to
the same table, tests/golden/make_golden.py) consume identical random numbers.

PINNING: tests/golden/make_golden.py runs the unmodified reference class on CPU and asserts
this module reproduces it; the resulting vectors are committed under tests/golden/.  The
`isaacgym.torch_utils` arithmetic underneath both is restated (oracle/torch_utils.py) and
unpinned by the reference itself -- see that file's header.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
import this module; the product path never does.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

from . import torch_utils as tu

# column map (kept numerically identical to deep-whole-body-control_b200/config.py; the
# oracle must not import the product package, so the constants are restated and a test
# asserts they agree)
RAND_GOAL_ORN, RAND_GOAL_SPH, RAND_CMD, RAND_PUSH = 0, 3, 33, 35
RAND_RST_DOF, RAND_RST_XY, RAND_RST_VEL, RAND_RST_CMD = 37, 57, 59, 65
RAND_RST_GOAL_ORN, RAND_RST_GOAL_SPH, RAND_TERRAIN, RAND_COLS = 67, 70, 100, 104

METRIC_NAMES = ["leg_energy_abs_sum", "tracking_lin_vel_x_l1", "tracking_ang_vel_yaw_exp",
                "tracking_ee_cart", "tracking_ee_sphere", "tracking_ee_orn", "leg_action_l2",
                "torque", "energy_square", "foot_contacts_z"]


def _u(rand, col, lo, hi):
    """torch_rand_float with the uniform taken from the table: (hi-lo)*r + lo, the span
    formed in float64 on the host exactly like the reference's python-float arithmetic."""
    return float(hi - lo) * rand[:, col] + float(lo)


class EnvOracle:
    """State + one `post_physics_step`.  `p` is any object exposing the attributes of
    WidowGo1Params (duck-typed; passed in by the tests)."""

    def __init__(self, p, s: SimpleNamespace):
        self.p = p
        self.s = s
        self.N = s.root_states_full.shape[0]
        self.leg_terms = p.active_terms("leg")
        self.arm_terms = p.active_terms("arm")
        self.extras = {"episode": {}}
        self.common_step_counter = 0
        if not hasattr(s, "episode_sums"):
            s.episode_sums = {k: torch.zeros(self.N) for k in p.sum_slots()}
            s.episode_metric_sums = {k: torch.zeros(self.N) for k in METRIC_NAMES}
        self._ig2r = torch.tensor(p.ig2raisim(), dtype=torch.long)
        self._ig2r18 = torch.tensor(p.ig2raisim(p.num_actions), dtype=torch.long)
        self._feet_perm = torch.tensor(p.feet_perm(), dtype=torch.long)
        if p.measure_heights:
            gx, gy = torch.meshgrid(torch.tensor(p.measured_points_x), torch.tensor(p.measured_points_y),
                                    indexing="ij")                                   # LR:783-785
            pts = torch.zeros(self.N, gx.numel(), 3)
            pts[:, :, 0] = gx.flatten()
            pts[:, :, 1] = gy.flatten()
            self.height_points = pts
        self.measured_heights = None

    # ---------------------------------------------------------------- views
    @property
    def root(self):
        return self.s.root_states_full[:, 0, :]

    @property
    def box(self):
        return self.s.root_states_full[:, 1, :]

    @property
    def dof_pos(self):
        return self.s.dof_state.view(self.N, -1, 2)[..., 0]

    @property
    def dof_vel(self):
        return self.s.dof_state.view(self.N, -1, 2)[..., 1]

    @property
    def contact_forces(self):
        return self.s.contact_forces_full[:, :-1, :]

    @property
    def ee_pos(self):
        return self.s.rigid_body_state[:, self.p.gripper_idx, :3]

    @property
    def ee_orn(self):
        return self.s.rigid_body_state[:, self.p.gripper_idx, 3:7]

    # ---------------------------------------------------------------- EE goal generator
    def _collision(self, start, goal):
        """WG:1337-1342 for every env (dense)."""
        p = self.p
        t = torch.linspace(0, 1, p.num_collision_check_samples)[None, :, None]          # [1,S,1]
        pts = tu.sphere2cart(torch.lerp(start[:, None, :], goal[:, None, :], t))           # [N,S,3]
        up = torch.tensor(p.collision_upper_limits, dtype=torch.float)
        lo = torch.tensor(p.collision_lower_limits, dtype=torch.float)
        inside = torch.all(pts < up, dim=-1) & torch.all(pts > lo, dim=-1)
        under = pts[..., 2] < p.underground_limit
        return inside.any(dim=1) | under.any(dim=1)

    def _resample_ee_goal(self, mask, rand, col_orn, col_sph, ranges):
        """WG:1316-1332 (+1303-1313) applied where `mask`; ranges = (l, p, y) [lo, hi] pairs."""
        s, p = self.s, self.p
        if not bool(mask.any()):
            return
        d = torch.stack([_u(rand, col_orn + i, p.final_delta_orn[i][0], p.final_delta_orn[i][1])
                         for i in range(3)], dim=-1)
        s.ee_goal_delta_orn_euler[mask] = d[mask]
        s.ee_goal_orn_euler[mask] = tu.torch_wrap_to_pi_minuspi(d + s.base_yaw_euler)[mask]
        s.ee_start_sphere[mask] = s.ee_goal_sphere[mask].clone()
        todo = mask.clone()
        for k in range(10):
            cand = torch.stack([_u(rand, col_sph + 3 * k + i, ranges[i][0], ranges[i][1])
                                for i in range(3)], dim=-1)
            s.ee_goal_sphere[todo] = cand[todo]
            todo = todo & self._collision(s.ee_start_sphere, s.ee_goal_sphere)
            if not bool(todo.any()):
                break
        s.ee_goal_cart[mask] = tu.sphere2cart(s.ee_goal_sphere)[mask]
        s.goal_timer[mask] = 0.0

    def _update_curr_ee_goal(self, rand, ranges):
        s = self.s                                                                       # WG:1344-1350
        t = torch.clip(s.goal_timer / s.traj_timesteps, 0, 1)
        s.curr_ee_goal_sphere[:] = torch.lerp(s.ee_start_sphere, s.ee_goal_sphere, t[:, None])
        s.curr_ee_goal_cart[:] = tu.sphere2cart(s.curr_ee_goal_sphere)
        s.goal_timer += 1
        self._resample_ee_goal(s.goal_timer > s.traj_total_timesteps, rand, RAND_GOAL_ORN, RAND_GOAL_SPH, ranges)

    # ---------------------------------------------------------------- commands / push
    def _resample_commands(self, mask, rand, col, rt):
        s, p = self.s, self.p                                                            # WG:831-843
        if not bool(mask.any()):
            return
        cx = _u(rand, col, rt.lin_vel_x[0], rt.lin_vel_x[1])
        cy = _u(rand, col + 1, rt.ang_vel_yaw[0], rt.ang_vel_yaw[1])
        keep = ((cx > p.lin_vel_x_clip) | (cy.abs() > p.ang_vel_yaw_clip)).float()
        new = torch.stack([cx * keep, torch.zeros_like(cx) * keep, cy * keep], dim=-1)
        s.commands[mask] = new[mask]

    def _push(self, rand):
        s, p = self.s, self.p                                                            # WG:804-814
        v = torch.stack([_u(rand, RAND_PUSH + i, -p.max_push_vel_xy, p.max_push_vel_xy) for i in range(2)], dim=-1)
        zero_cmd = (s.commands.sum(dim=1) == 0).unsqueeze(-1)
        self.root[:, 7:9] = torch.where(zero_cmd, v * 2.5, v)

    # ---------------------------------------------------------------- heights (LR:793-829)
    def _get_heights(self):
        s, p = self.s, self.p
        q = self.root[:, 3:7].clone()
        q[:, :2] = 0.0                                                                   # utils/math.py:38-42
        q = tu.normalize(q)
        npts = self.height_points.shape[1]
        pts = tu.quat_apply(q.repeat(1, npts), self.height_points) + self.root[:, :3].unsqueeze(1)
        pts = pts + p.border_size
        pts = (pts / p.horizontal_scale).long()
        px = torch.clip(pts[:, :, 0].reshape(-1), 0, s.height_samples.shape[0] - 2)
        py = torch.clip(pts[:, :, 1].reshape(-1), 0, s.height_samples.shape[1] - 2)
        h = torch.min(torch.min(s.height_samples[px, py], s.height_samples[px + 1, py]), s.height_samples[px, py + 1])
        return h.view(self.N, -1) * p.vertical_scale

    # ---------------------------------------------------------------- termination (WG:937-963)
    def _check_termination(self):
        s, p = self.s, self.p
        idx = torch.tensor(p.termination_contact_indices, dtype=torch.long)
        contact = torch.any(torch.norm(self.contact_forces[:, idx, :], dim=-1) > 1.0, dim=1)
        r, pt, _ = tu.euler_from_quat(self.root[:, 3:7])
        g = s.curr_ee_goal_cart if p.command_mode == "cart" else s.curr_ee_goal_sphere
        r_bad = ((r > p.term_roll) & (g[:, 2] >= 0)) | ((r < -p.term_roll) & (g[:, 2] <= 0))
        p_bad = ((pt > p.term_pitch) & (g[:, 1] >= 0)) | ((pt < -p.term_pitch) & (g[:, 1] <= 0))
        z_bad = self.root[:, 2] < p.term_z
        s.time_out_buf = s.episode_length_buf > p.max_episode_length
        s.reset_buf = contact | r_bad | p_bad | z_bad | s.time_out_buf

    # ---------------------------------------------------------------- reward terms
    def _term(self, name):
        s, p, m = self.s, self.p, self.s.episode_metric_sums
        tq, dv, act, cmd = s.torques, self.dof_vel, s.actions, s.commands
        feet = torch.tensor(p.feet_indices, dtype=torch.long)
        cf = self.contact_forces
        if name == "energy_square":                                                      # WG:1466-1469
            e = torch.sum(torch.square(tq[:, :12] * dv[:, :12]), dim=1)
            m["energy_square"] += e
            return e
        if name == "foot_contacts_z":                                                    # WG:1455-1458
            f = torch.square(s.force_sensor[:, :, 2]).sum(dim=-1)
            m["foot_contacts_z"] += f
            return f
        if name == "hip_action_l2":                                                      # WG:1379-1382
            a = torch.sum(act[:, [0, 3, 6, 9]] ** 2, dim=1)
            m["leg_action_l2"] += a
            return a
        if name == "leg_action_l2":                                                      # WG:1405-1408
            a = torch.sum(act[:, :12] ** 2, dim=1)
            m["leg_action_l2"] += a
            return a
        if name == "survive":                                                            # WG:1452-1453
            return torch.ones(self.N)
        if name == "tracking_ang_vel_yaw_exp":                                           # WG:1441-1444
            e = torch.abs(cmd[:, 2] - s.base_ang_vel[:, 2])
            m["tracking_ang_vel_yaw_exp"] += e
            return torch.exp(-e / p.tracking_sigma)
        if name == "tracking_ang_vel_yaw_l1":                                            # WG:1437-1439
            e = torch.abs(cmd[:, 2] - s.base_ang_vel[:, 2])
            return -e + torch.abs(cmd[:, 2])
        if name == "tracking_lin_vel_x_l1":                                              # WG:1427-1430
            e = torch.abs(cmd[:, 0] - s.base_lin_vel[:, 0])
            m["tracking_lin_vel_x_l1"] += e
            return -e + torch.abs(cmd[:, 0])
        if name == "tracking_lin_vel_x_exp":                                             # WG:1432-1435
            e = torch.abs(cmd[:, 0] - s.base_lin_vel[:, 0])
            m["tracking_lin_vel_x_l1"] += e
            return torch.exp(-e / p.tracking_sigma)
        if name == "tracking_lin_vel_y_l2":                                              # WG:1446-1447
            return (cmd[:, 1] - s.base_lin_vel[:, 1]) ** 2
        if name == "tracking_lin_vel_z_l2":                                              # WG:1449-1450
            return (cmd[:, 2] - s.base_lin_vel[:, 2]) ** 2
        if name == "tracking_lin_vel":                                                   # WG:1422-1425
            e = torch.sum(torch.square(cmd[:, :2] - s.base_lin_vel[:, :2]), dim=1)
            return torch.exp(-e / p.tracking_sigma)
        if name == "tracking_ang_vel":                                                   # LR:886-889
            e = torch.square(cmd[:, 2] - s.base_ang_vel[:, 2])
            return torch.exp(-e / p.tracking_sigma)
        if name == "torques":                                                            # WG:1460-1464
            t = torch.sum(torch.square(tq), dim=1)
            m["torque"] += t
            return t
        if name == "leg_energy_abs_sum":                                                 # WG:1396-1399
            e = torch.sum(torch.abs(tq[:, :12] * dv[:, :12]), dim=1)
            m["leg_energy_abs_sum"] += e
            return e
        if name == "leg_energy_sum_abs":                                                 # WG:1401-1403
            return torch.abs(torch.sum(tq[:, :12] * dv[:, :12], dim=1))
        if name == "leg_energy":                                                         # WG:1410-1412
            return torch.sum(tq[:, :12] * dv[:, :12], dim=1)
        if name == "arm_energy_abs_sum":                                                 # WG:1414-1415
            return torch.sum(torch.abs(tq[:, 12:-2] * dv[:, 12:-2]), dim=1)
        if name == "tracking_ee_sphere":                                                 # WG:1352-1358
            off = torch.cat([self.root[:, :2], torch.full((self.N, 1), p.z_invariant_offset)], dim=1)
            loc = tu.quat_rotate_inverse(s.base_yaw_quat, self.ee_pos - off)
            scale = torch.tensor(p.sphere_error_scale, dtype=torch.float)
            e = torch.sum(torch.abs(tu.cart2sphere(loc) - s.curr_ee_goal_sphere) * scale, dim=1)
            m["tracking_ee_sphere"] += e
            return torch.exp(-e / p.tracking_ee_sigma)
        if name == "tracking_ee_cart":                                                   # WG:1360-1366
            off = torch.cat([self.root[:, :2], torch.full((self.N, 1), p.z_invariant_offset)], dim=1)
            tgt = off + tu.quat_apply(s.base_yaw_quat, s.curr_ee_goal_cart)
            e = torch.sum(torch.abs(self.ee_pos - tgt), dim=1)
            m["tracking_ee_cart"] += e
            return torch.exp(-e / p.tracking_ee_sigma)
        if name in ("tracking_ee_orn", "tracking_ee_orn_ry"):                            # WG:1368-1394
            eul = torch.stack(tu.euler_from_quat(self.ee_orn), dim=-1)
            scale = torch.tensor(p.orn_error_scale, dtype=torch.float)
            d = tu.torch_wrap_to_pi_minuspi(s.ee_goal_orn_euler - eul)
            if name == "tracking_ee_orn":
                e = torch.sum(torch.abs(d) * scale, dim=1)
            else:
                e = torch.sum(torch.abs((d * scale)[:, [0, 2]]), dim=1)
                m["tracking_ee_orn"] += e
            return torch.exp(-e / p.tracking_ee_sigma)
        if name == "lin_vel_z":                                                          # LR:832-834
            return torch.square(s.base_lin_vel[:, 2])
        if name == "ang_vel_xy":                                                         # LR:836-838
            return torch.sum(torch.square(s.base_ang_vel[:, :2]), dim=1)
        if name == "base_height":                                                        # LR:844-847
            bh = torch.mean(self.root[:, 2].unsqueeze(1) - self.measured_heights, dim=1)
            return torch.square(bh - p.base_height_target)
        if name == "dof_vel":                                                            # LR:853-855
            return torch.sum(torch.square(dv), dim=1)
        if name == "dof_acc":                                                            # LR:857-859
            return torch.sum(torch.square((s.last_dof_vel - dv) / p.dt), dim=1)
        if name == "action_rate":                                                        # LR:861-863
            return torch.sum(torch.square(s.last_actions - act), dim=1)
        if name == "collision":                                                          # LR:865-867
            idx = torch.tensor(p.penalized_contact_indices, dtype=torch.long)
            return torch.sum(1.0 * (torch.norm(cf[:, idx, :], dim=-1) > 0.1), dim=1)
        if name == "termination":                                                        # LR:869-871
            return (s.reset_buf & ~s.time_out_buf).float()
        if name == "dof_pos_limits":                                                     # LR:873-877
            lim = torch.tensor(p.dof_pos_limits, dtype=torch.float)
            out = -(self.dof_pos - lim[:, 0]).clip(max=0.0)
            out = out + (self.dof_pos - lim[:, 1]).clip(min=0.0)
            return torch.sum(out, dim=1)
        if name == "dof_vel_limits":                                                     # LR:879-882
            lim = torch.tensor(p.dof_vel_limits, dtype=torch.float)
            return torch.sum((torch.abs(dv) - lim * p.soft_dof_vel_limit).clip(min=0.0, max=1.0), dim=1)
        if name == "torque_limits":                                                      # LR:884-886
            lim = torch.tensor(p.torque_limits, dtype=torch.float)
            return torch.sum((torch.abs(tq) - lim * p.soft_torque_limit).clip(min=0.0), dim=1)
        if name == "feet_air_time":                                                      # LR:896-908
            contact = cf[:, feet, 2] > 1.0
            filt = contact | s.last_contacts
            s.last_contacts = contact
            first = (s.feet_air_time > 0.0) * filt
            s.feet_air_time += p.dt
            r = torch.sum((s.feet_air_time - 0.5) * first, dim=1)
            r = r * (torch.norm(cmd[:, :2], dim=1) > 0.1)
            s.feet_air_time *= ~filt
            return r
        if name == "stumble":                                                            # LR:910-913
            return torch.any(torch.norm(cf[:, feet, :2], dim=2) > 5 * torch.abs(cf[:, feet, 2]), dim=1).float()
        if name == "stand_still":                                                        # LR:915-917
            return torch.sum(torch.abs(self.dof_pos - torch.tensor(p.default_dof_pos)), dim=1) * \
                (torch.norm(cmd[:, :2], dim=1) < 0.1)
        if name == "feet_contact_forces":                                                # LR:919-921
            return torch.sum((torch.norm(cf[:, feet, :], dim=-1) - p.max_contact_force).clip(min=0.0), dim=1)
        raise KeyError(name)

    def _compute_reward(self, leg_scales, arm_scales):
        """WG:170-205: per channel, alphabetical accumulation, optional clip, termination, /100."""
        s, p = self.s, self.p
        out = []
        for terms, scales in ((self.leg_terms, leg_scales), (self.arm_terms, arm_scales)):
            buf = torch.zeros(self.N)
            for name in terms:
                rew = self._term(name) * scales[name]
                buf += rew
                s.episode_sums[name] += rew
            if p.only_positive_rewards:
                buf = torch.clip(buf, min=0.0)
            if scales.get("termination", 0) != 0 and "termination" in s.episode_sums:
                rew = self._term("termination") * scales["termination"]
                buf += rew
                s.episode_sums["termination"] += rew
            out.append(buf / 100)
        s.rew_buf, s.arm_rew_buf = out

    # ---------------------------------------------------------------- reset (WG:695-754)
    def _terrain_curriculum(self, mask, rand):
        s, p = self.s, self.p                                                            # LR:421-441
        dist = torch.norm(self.root[:, :2] - s.env_origins[:, :2], dim=1)
        up = dist > p.terrain_env_length / 2
        down = (dist < torch.norm(s.commands[:, :2], dim=1) * p.max_episode_length_s * 0.5) * ~up
        lvl = s.terrain_levels + 1 * up - 1 * down
        rnd = torch.clamp((rand[:, RAND_TERRAIN] * p.max_terrain_level).long(), max=p.max_terrain_level - 1)
        lvl = torch.where(lvl >= p.max_terrain_level, rnd, torch.clip(lvl, 0))
        s.terrain_levels[mask] = lvl[mask]
        s.env_origins[mask] = s.terrain_origins[s.terrain_levels, s.terrain_types][mask]

    def _reset(self, rand, rt, goal_ranges):
        s, p = self.s, self.p
        mask = s.reset_buf.clone()
        self.reset_count = int(mask.sum())
        if self.reset_count == 0:
            return
        if p.terrain_curriculum:
            self._terrain_curriculum(mask, rand)
        # _reset_dofs WG:816-828
        default = torch.tensor(p.default_dof_pos, dtype=torch.float)
        u = torch.stack([_u(rand, RAND_RST_DOF + i, 0.8, 1.2) for i in range(p.num_dofs)], dim=-1)
        self.dof_pos[mask] = (default * u)[mask]
        self.dof_vel[mask] = 0.0
        # _reset_root_states WG:757-788
        base = torch.tensor(p.base_init_state, dtype=torch.float).repeat(self.N, 1)
        base[:, :3] += s.env_origins
        pr = p.origin_perturb_range
        base[:, :2] += torch.stack([_u(rand, RAND_RST_XY + i, -pr, pr) for i in range(2)], dim=-1)
        vr = p.init_vel_perturb_range
        base[:, 7:13] = torch.stack([_u(rand, RAND_RST_VEL + i, -vr, vr) for i in range(6)], dim=-1)
        self.root[mask] = base[mask]
        bx = self.box.clone()
        bx[:, 0] = p.box_env_origins_x
        bx[:, 1] = self.root[:, 1] + s.box_env_origins_delta_y
        bx[:, 2] = p.box_env_origins_z
        self.box[mask, :3] = bx[mask, :3]
        # commands only for timed-out envs (WG:723-727); EE goal for every reset env
        self._resample_commands(s.time_out_buf.clone(), rand, RAND_RST_CMD, rt)
        self._resample_ee_goal(mask, rand, RAND_RST_GOAL_ORN, RAND_RST_GOAL_SPH, goal_ranges)
        s.last_actions[mask] = 0.0
        s.last_dof_vel[mask] = 0.0
        s.feet_air_time[mask] = 0.0
        s.episode_length_buf[mask] = 0
        s.obs_history_buf[mask] = 0.0
        s.action_history_buf[mask] = 0.0
        s.goal_timer[mask] = 0.0
        ep = {}
        for k in s.episode_sums:                                                         # WG:743-750
            ep["rew_" + k] = torch.mean(s.episode_sums[k][mask]) / p.max_episode_length_s
            s.episode_sums[k][mask] = 0.0
        for k in s.episode_metric_sums:
            ep["metric_" + k] = torch.mean(s.episode_metric_sums[k][mask]) / p.max_episode_length_s
            s.episode_metric_sums[k][mask] = 0.0
        self.extras["episode"] = ep
        self.extras["time_outs"] = s.time_out_buf

    # ---------------------------------------------------------------- observations (WG:966-1001)
    def _compute_observations(self):
        s, p = self.s, self.p
        wrapped = self.dof_pos.clone()
        wrapped[:, -8] = tu.torch_wrap_to_pi_minuspi(wrapped[:, -8])
        s.dof_pos_wrapped = wrapped
        r, pt, _ = tu.euler_from_quat(self.root[:, 3:7])
        contacts = (s.force_sensor.norm(dim=-1) > 1.5)                                   # WG:1090-1098
        default = torch.tensor(p.default_dof_pos, dtype=torch.float)
        cscale = torch.tensor([p.obs_scale_lin_vel, p.obs_scale_lin_vel, p.obs_scale_ang_vel])
        goal = s.curr_ee_goal_cart if p.command_mode == "cart" else s.curr_ee_goal_sphere
        prop = torch.cat((
            torch.stack([r, pt], dim=-1),
            s.base_ang_vel * p.obs_scale_ang_vel,
            ((wrapped - default) * p.obs_scale_dof_pos)[:, self._ig2r],
            (self.dof_vel * p.obs_scale_dof_vel)[:, self._ig2r],
            s.action_history_buf[:, -1][:, self._ig2r18],
            contacts[:, self._feet_perm],
            s.commands[:, :3] * cscale,
            goal,
            s.ee_goal_delta_orn_euler), dim=-1)
        priv = torch.cat((s.mass_params, s.friction, s.motor_strength - 1), dim=-1)
        s.obs_buf = torch.cat([prop, priv, s.obs_history_buf.view(self.N, -1)], dim=-1)
        s.obs_history_buf = torch.where(
            (s.episode_length_buf <= 1)[:, None, None],
            torch.stack([prop] * p.history_len, dim=1),
            torch.cat([s.obs_history_buf[:, 1:], prop.unsqueeze(1)], dim=1))
        s.prop = prop

    # ---------------------------------------------------------------- driver (WG:865-915)
    def post_physics_step(self, rand, rt, clip_obs=True):
        """rt: runtime namespace with lin_vel_x, ang_vel_yaw, goal_l, goal_p, goal_y ([lo,hi],
        float64) and leg_scales / arm_scales dicts (the curriculum outputs, WG:678-692)."""
        s, p = self.s, self.p
        s.episode_length_buf += 1
        self.common_step_counter += 1
        q = self.root[:, 3:7]
        s.base_lin_vel[:] = tu.quat_rotate_inverse(q, self.root[:, 7:10])
        s.base_ang_vel[:] = tu.quat_rotate_inverse(q, self.root[:, 10:13])
        yaw = tu.euler_from_quat(q)[2]
        s.base_yaw_euler[:] = torch.cat([torch.zeros(self.N, 2), yaw.view(-1, 1)], dim=1)
        s.base_yaw_quat[:] = tu.quat_from_euler_xyz(torch.tensor(0), torch.tensor(0), yaw)
        goal_ranges = (rt.goal_l, rt.goal_p, rt.goal_y)
        self._update_curr_ee_goal(rand, goal_ranges)
        # _post_physics_step_callback WG:917-935
        self._resample_commands(s.episode_length_buf % p.resample_interval == 0, rand, RAND_CMD, rt)
        if p.measure_heights:
            self.measured_heights = self._get_heights()
        self.pushed = bool(p.push_robots and (self.common_step_counter % p.push_interval == 0))
        if self.pushed:
            self._push(rand)
        self._check_termination()
        self._compute_reward(rt.leg_scales, rt.arm_scales)
        self._reset(rand, rt, goal_ranges)
        self._compute_observations()
        s.last_actions[:] = s.actions
        s.last_dof_vel[:] = self.dof_vel
        s.last_root_vel[:] = self.root[:, 7:13]
        if clip_obs:                                                                     # WG:1195-1196
            s.obs_buf = torch.clip(s.obs_buf, -p.clip_observations, p.clip_observations)
        return s.obs_buf, s.rew_buf, s.arm_rew_buf, s.reset_buf, self.extras


def heights_obs(root_z, measured_heights, scale):
    """Perceptive obs formula of the base class (LR:221-223)."""
    return torch.clip(root_z.unsqueeze(1) - 0.5 - measured_heights, -1, 1.0) * scale


def compute_torques(actions, dof_pos, dof_vel, motor_strength, p_gains, d_gains, action_scale, default_dof_pos, torque_limits, wrap_col=-8):
    """`WidowGo1._compute_torques` (WG:1262-1295) with adaptive_arm_gains = torque_supervision = False (WGC:168,173).

    actions / motor_strength [N, n_act] (delayed actions, Isaac Gym order), dof_pos / dof_vel [N, n_dof], gains / scale [n_act],
    default_dof_pos / torque_limits [n_dof].  Returns torques [N, n_dof] (zero for the non-driven gripper DOFs, WG:1291).
    The reference wraps column -8 of the n_act-wide position tensor (WG:1279); that is DOF n_act-8, restated as written."""
    from .torch_utils import torch_wrap_to_pi_minuspi
    n_act = actions.shape[1]
    scaled = actions * motor_strength * action_scale                                        # WG:1276
    q = dof_pos[:, :n_act].clone()                                                          # WG:1278
    q[:, wrap_col] = torch_wrap_to_pi_minuspi(q[:, wrap_col])                               # WG:1279
    tau = p_gains * (scaled + default_dof_pos[:n_act] - q) - d_gains * dof_vel[:, :n_act]   # WG:1281
    tau = torch.cat([tau, torch.zeros(actions.shape[0], dof_pos.shape[1] - n_act)], dim=-1)  # WG:1291
    return torch.clip(tau, -torque_limits, torque_limits)                                   # WG:1295
