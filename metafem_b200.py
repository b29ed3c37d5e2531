"""Import shim: the package directory is named ``metafem.jl_b200`` (not a valid module name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "metafem.jl_b200")
_name = "metafem_jl_b200"
if _name not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_name, os.path.join(_dir, "__init__.py"),
                                                   submodule_search_locations=[_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
pkg = sys.modules[_name]
globals().update({k: getattr(pkg, k) for k in dir(pkg) if not k.startswith("__")})
