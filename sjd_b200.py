"""Import alias: the package directory is named ``accelerating-t2i-ar-with-sjd_b200`` (not a valid Python
identifier), so ``import sjd_b200`` loads it from that directory under this name."""
import importlib.util as _u
import sys as _sys
from pathlib import Path as _P

_dir = _P(__file__).resolve().parent / "accelerating-t2i-ar-with-sjd_b200"
_spec = _u.spec_from_file_location("sjd_b200", _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = _u.module_from_spec(_spec)
_sys.modules["sjd_b200"] = _mod
_spec.loader.exec_module(_mod)
