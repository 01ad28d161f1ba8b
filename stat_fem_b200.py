"""Import shim: the package lives in ``stat-fem_b200/`` (a name Python cannot import directly); this module
loads it under the importable name ``stat_fem_b200`` so that ``import stat_fem_b200 as stat_fem`` works
from the repository root."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "stat-fem_b200")
_spec = _ilu.spec_from_file_location("stat_fem_b200", _os.path.join(_pkg_dir, "__init__.py"),
                                     submodule_search_locations=[_pkg_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["stat_fem_b200"] = _mod
_spec.loader.exec_module(_mod)
