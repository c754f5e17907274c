"""Import alias: ``import vlsat_b200`` loads the package that lives in ``cvpr2023-vlsat_b200/``.

The product directory carries the reference's name (hyphen included), which is not a legal Python
identifier, so this one-file loader registers it under an importable name.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_pkg_dir = _os.path.join(_here, "cvpr2023-vlsat_b200")
_spec = _ilu.spec_from_file_location(
    "vlsat_b200", _os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["vlsat_b200"] = _mod
_spec.loader.exec_module(_mod)
