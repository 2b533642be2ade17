"""TEST INFRASTRUCTURE ONLY -- CPU restatement of ``isaacgym.torch_utils``.

The reference (`legged_gym/legged_gym/envs/widowGo1/widowGo1.py:35`,
`envs/base/legged_robot.py:37`, `utils/math.py:34`) star-imports this module from
the closed-source pip package ``isaacgym`` (Preview 3/4, no version pin:
`legged_gym/setup.py:12`).  The package is not vendored under /root/reference and is
not installed, so the arithmetic is restated here from its published definitions.

Two groups:

* stock Isaac Gym helpers (public, stable definitions): ``quat_rotate_inverse``,
  ``quat_apply``, ``quat_from_euler_xyz``, ``normalize``, ``torch_rand_float``,
  ``to_torch``, ``get_axis_params``, ``quat_mul``, ``quat_conjugate``;
* helpers the reference authors added to their private copy and that exist nowhere
  in the tree: ``euler_from_quat``, ``sphere2cart``, ``cart2sphere``,
  ``torch_wrap_to_pi_minuspi``, ``torch_rand_sign``, ``orientation_error``.  Their
  definitions are INFERRED from call sites (`widowGo1.py:855-863` holds the commented
  out in-tree spherical convention; `widowGo1.py:942-946` fixes roll/pitch/yaw order).
  Parity at this boundary is therefore *unpinned by the reference itself*: this file
  is the single source of truth shared by the oracle, by the fake ``isaacgym`` package
  used to execute the reference's own Python (tests/fakes), and by the CUDA kernels.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path never does.
"""
import math

import numpy as np
import torch

__all__ = [
    "quat_rotate_inverse", "quat_rotate", "quat_apply", "quat_from_euler_xyz", "normalize",
    "torch_rand_float", "to_torch", "get_axis_params", "quat_mul", "quat_conjugate",
    "euler_from_quat", "sphere2cart", "cart2sphere", "torch_wrap_to_pi_minuspi",
    "torch_rand_sign", "orientation_error",
]


def to_torch(x, dtype=torch.float, device="cpu", requires_grad=False):
    return torch.tensor(x, dtype=dtype, device=device, requires_grad=requires_grad)


def get_axis_params(value, axis_idx, x_value=0.0, dtype=float, n_dims=3):
    axis = np.zeros((n_dims,))
    axis[axis_idx] = 1.0
    out = np.where(axis == 1.0, value, axis)
    out[0] = x_value
    return list(out.astype(dtype))


def normalize(x, eps: float = 1e-9):
    return x / x.norm(p=2, dim=-1).clamp(min=eps, max=None).unsqueeze(-1)


def quat_mul(a, b):
    shape = a.shape
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 4)
    ax, ay, az, aw = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    bx, by, bz, bw = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    x = aw * bx + ax * bw + ay * bz - az * by
    y = aw * by - ax * bz + ay * bw + az * bx
    z = aw * bz + ax * by - ay * bx + az * bw
    w = aw * bw - ax * bx - ay * by - az * bz
    return torch.stack([x, y, z, w], dim=-1).view(shape)


def quat_conjugate(a):
    shape = a.shape
    a = a.reshape(-1, 4)
    return torch.cat((-a[:, :3], a[:, -1:]), dim=-1).view(shape)


def quat_apply(q, v):
    """v + 2 w (q_v x v) + 2 q_v x (q_v x v); q is xyzw."""
    shape = v.shape
    q = q.reshape(-1, 4)
    v = v.reshape(-1, 3)
    qv = q[:, :3]
    t = torch.cross(qv, v, dim=-1) * 2
    return (v + q[:, 3:] * t + torch.cross(qv, t, dim=-1)).view(shape)


def quat_rotate(q, v):
    qw = q[:, -1]
    qv = q[:, :3]
    a = v * (2.0 * qw ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(qv, v, dim=-1) * qw.unsqueeze(-1) * 2.0
    c = qv * (qv * v).sum(dim=-1, keepdim=True) * 2.0
    return a + b + c


def quat_rotate_inverse(q, v):
    """v (2w^2-1) - 2 w (q_v x v) + 2 q_v (q_v . v); q is xyzw."""
    qw = q[:, -1]
    qv = q[:, :3]
    a = v * (2.0 * qw ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(qv, v, dim=-1) * qw.unsqueeze(-1) * 2.0
    c = qv * (qv * v).sum(dim=-1, keepdim=True) * 2.0
    return a - b + c


def quat_from_euler_xyz(roll, pitch, yaw):
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    qw = cy * cr * cp + sy * sr * sp
    qx = cy * sr * cp - sy * cr * sp
    qy = cy * cr * sp + sy * sr * cp
    qz = sy * cr * cp - cy * sr * sp
    return torch.stack([qx, qy, qz, qw], dim=-1)


def torch_rand_float(lower, upper, shape, device):
    return (upper - lower) * torch.rand(*shape, device=device) + lower


# --------------------------------------------------------------------------------------
# private additions of the reference authors (inferred; see module docstring)
# --------------------------------------------------------------------------------------

def euler_from_quat(q):
    """xyzw quaternion -> (roll, pitch, yaw), each [N] (unpacked at widowGo1.py:942)."""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    roll = torch.atan2(2.0 * (w * x + y * z), 1.0 - 2.0 * (x * x + y * y))
    pitch = torch.asin(torch.clip(2.0 * (w * y - z * x), -1.0, 1.0))
    yaw = torch.atan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z))
    return roll, pitch, yaw


def sphere2cart(s):
    """(l, pitch, yaw) -> (x, y, z); convention of the commented block widowGo1.py:855-863."""
    l, p, y = s[..., 0], s[..., 1], s[..., 2]
    proj = l * torch.cos(p)
    return torch.stack([proj * torch.cos(y), proj * torch.sin(y), l * torch.sin(p)], dim=-1)


def cart2sphere(c):
    x, y, z = c[..., 0], c[..., 1], c[..., 2]
    l = torch.sqrt(x * x + y * y + z * z)
    return torch.stack([l, torch.asin(z / l), torch.atan2(y, x)], dim=-1)


def torch_wrap_to_pi_minuspi(a):
    return torch.remainder(a + math.pi, 2.0 * math.pi) - math.pi


def torch_rand_sign(shape, device):
    return 2.0 * torch.randint(0, 2, shape, device=device).float() - 1.0


def orientation_error(desired, current):
    cc = quat_conjugate(current)
    q_r = quat_mul(desired, cc)
    return q_r[:, 0:3] * torch.sign(q_r[:, 3]).unsqueeze(-1)
