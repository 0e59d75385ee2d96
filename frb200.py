"""Import shim: the package directory is named ``fluxreconstruction.jl_b200`` (not a valid
Python identifier), so it is loaded here under the module name ``frb200``.

    import frb200 as FR
    ps = FR.FRPSpace2D(0.0, 1.0, 64, 0.0, 1.0, 64, 3, 1, 1)
"""
import importlib.util as _u
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "fluxreconstruction.jl_b200")
_spec = _u.spec_from_file_location(
    "frb200", _os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
_mod = _u.module_from_spec(_spec)
_sys.modules["frb200"] = _mod
_spec.loader.exec_module(_mod)
