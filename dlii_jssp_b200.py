"""Import shim: the package lives in ``dlii-jssp_b200/`` (a directory name Python cannot import); this module loads it
under the importable name ``dlii_jssp_b200`` and replaces itself in ``sys.modules``."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "dlii-jssp_b200")
_spec = _u.spec_from_file_location("dlii_jssp_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["dlii_jssp_b200"] = _mod
_spec.loader.exec_module(_mod)
