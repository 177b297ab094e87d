"""Import shim: makes the in-tree directory `legenddsp.jl_b200/` importable as the module `legenddsp.jl_b200`
(a directory name containing a dot cannot be found by the normal import machinery)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "legenddsp.jl_b200")


def _load():
    name = "legenddsp.jl_b200"
    if name in _sys.modules:
        return _sys.modules[name]
    spec = _ilu.spec_from_file_location(name, _os.path.join(_PKG_DIR, "__init__.py"),
                                        submodule_search_locations=[_PKG_DIR])
    mod = _ilu.module_from_spec(spec)
    _sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


jl_b200 = _load()
