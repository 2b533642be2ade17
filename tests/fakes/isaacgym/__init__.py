"""Fake ``isaacgym`` package -- TEST INFRASTRUCTURE ONLY.

Lets tests/golden/make_golden.py import the reference's own Python
(`legged_gym/legged_gym/envs/widowGo1/widowGo1.py`) in this container, where the
closed-source Isaac Gym wheel is absent.  Physics entry points are inert stubs; the
only real arithmetic is ``isaacgym.torch_utils`` which re-exports the restatement in
``oracle/torch_utils.py`` (single source of truth).
"""
import sys
import types


class _Anything:
    """Callable/attribute sink: every attribute is another sink, every call returns one."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())

    def __bool__(self):
        return False


def _stub_module(name):
    mod = types.ModuleType(name)
    mod.__getattr__ = lambda attr, _n=name: (_ for _ in ()).throw(AttributeError(attr)) \
        if (attr.startswith("__") and attr.endswith("__")) else _Anything()
    sys.modules[name] = mod
    return mod


gymapi = _stub_module("isaacgym.gymapi")
gymtorch = _stub_module("isaacgym.gymtorch")
gymutil = _stub_module("isaacgym.gymutil")
terrain_utils = _stub_module("isaacgym.terrain_utils")
from . import torch_utils  # noqa: E402,F401
