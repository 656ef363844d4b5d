"""Import shim: makes ``import tnrkit.jl_b200`` resolve to the package directory
``tnrkit.jl_b200/`` at the repository root (a directory name with a dot cannot be
imported directly)."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(os.path.dirname(_here), "tnrkit.jl_b200")
if "tnrkit.jl_b200" not in sys.modules:
    _spec = importlib.util.spec_from_file_location(
        "tnrkit.jl_b200", os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg]
    )
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules["tnrkit.jl_b200"] = _mod
    _spec.loader.exec_module(_mod)
jl_b200 = sys.modules["tnrkit.jl_b200"]
