"""Import shim: the package directory is named after the reference repo (`amira-rust-asr-server_b200/`), which is not
a valid Python identifier; this module loads it under the name `amira_b200`."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "amira-rust-asr-server_b200")
_spec = _u.spec_from_file_location("amira_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["amira_b200"] = _mod
_spec.loader.exec_module(_mod)
