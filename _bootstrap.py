"""Loads the product package directory `finitediff.jl_b200/` (whose name contains a dot, so
it cannot be imported by name) under the importable alias `finitediff_jl_b200`."""
import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
_ALIAS = "finitediff_jl_b200"


def load_package():
    if _ALIAS in sys.modules:
        return sys.modules[_ALIAS]
    pkg_dir = ROOT / "finitediff.jl_b200"
    spec = importlib.util.spec_from_file_location(_ALIAS, pkg_dir / "__init__.py",
                                                  submodule_search_locations=[str(pkg_dir)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_ALIAS] = mod
    spec.loader.exec_module(mod)
    return mod
