"""Import alias: ``import dwbc_b200`` loads the package directory
``deep-whole-body-control_b200/`` (whose name is not a valid Python identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "deep-whole-body-control_b200")
_spec = importlib.util.spec_from_file_location(
    "dwbc_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dwbc_b200"] = _mod
_spec.loader.exec_module(_mod)
