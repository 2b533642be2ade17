"""Fake ``isaacgym.torch_utils``: re-export of oracle/torch_utils.py (see its docstring)."""
import os
import sys

_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from oracle.torch_utils import *  # noqa: E402,F401,F403
from oracle.torch_utils import __all__  # noqa: E402,F401
