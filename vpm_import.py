"""Import helper: the package directory is named `flowvpm.jl_b200` (with a dot),
which the import statement cannot spell, so it is registered as
`flowvpm_jl_b200`.  Usage:  `from vpm_import import vpm`."""
import importlib.util
import os
import sys

_NAME = "flowvpm_jl_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "flowvpm.jl_b200")


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


def load_build():
    """the build helper alone (does not import the package or touch the GPU)"""
    spec = importlib.util.spec_from_file_location(_NAME + "_build", os.path.join(_DIR, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Lazy:
    def __getattr__(self, name):
        return getattr(load(), name)


vpm = _Lazy()
