"""Importable alias for the package directory ``clarabel.rs_b200/`` (the dot in
the directory name makes it unimportable by the normal machinery)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "clarabel.rs_b200")
_spec = importlib.util.spec_from_file_location(
    "clarabel_rs_b200_pkg", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["clarabel_rs_b200_pkg"] = _mod
_spec.loader.exec_module(_mod)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
pkg = _mod
