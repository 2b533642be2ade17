"""dwbc_b200 -- Blackwell-native (sm_100a) hot path of MarkFzp/Deep-Whole-Body-Control.

Scope: the widowGo1 post-physics step (obs / reward / termination /
reset / history, height scan) and the rsl_rl PPO update path (GAE, ActorCritic forward and
backward, PPO loss, clip + Adam, gradient all-reduce), as hand-written CUDA kernels behind a
C ABI (include/dwbc.h) with a host-side mirror of the reference's Python surfaces.

Importing the package never touches CUDA; the native library is loaded on first use by
`_lib.lib()` and raises if it is missing (no CPU fallback exists on the product path).
"""
from .config import WidowGo1Params, CommandCurriculum  # noqa: F401

__all__ = ["WidowGo1Params", "CommandCurriculum"]
