"""Import shim: ``import sdpsr_b200`` loads the package that lives in the
directory ``sdpsymmetryreduction.jl_b200/`` (its name contains a dot, so Python's
import system cannot find it by name)."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_here, "sdpsymmetryreduction.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "sdpsr_b200", os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sdpsr_b200"] = _mod
_spec.loader.exec_module(_mod)
