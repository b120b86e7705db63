"""Import shim: exposes the package directory ``learning-from-sparse-demonstrations_b200/``
(not a valid Python identifier) under the importable name ``lfsd_b200``."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                         "learning-from-sparse-demonstrations_b200")
_spec = _ilu.spec_from_file_location(
    "lfsd_b200", _os.path.join(_pkg_dir, "__init__.py"),
    submodule_search_locations=[_pkg_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["lfsd_b200"] = _mod
_spec.loader.exec_module(_mod)
