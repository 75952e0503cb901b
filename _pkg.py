"""Loader for the package directory `gemotion.jl_b200/` (its name is not a Python identifier)."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_NAME = "gemotion_jl_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_ROOT, "gemotion.jl_b200", "__init__.py"),
        submodule_search_locations=[os.path.join(_ROOT, "gemotion.jl_b200")])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
gm = sys.modules[_NAME]
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
