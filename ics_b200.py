"""Import shim: the package directory is named ``image-classification-system_b200`` (not a
valid Python identifier), so ``import ics_b200`` loads it from that directory under this name.

    import ics_b200
    from ics_b200.services.webdav_sync import WebDAVSync
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "image-classification-system_b200")
_spec = _ilu.spec_from_file_location(
    "ics_b200", _os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["ics_b200"] = _mod
_spec.loader.exec_module(_mod)
