"""Loader: the package directory is `crystal-packing_b200/` (hyphenated, as the project is named),
which Python cannot import by name.  `import crystal_packing_b200` loads it from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "crystal-packing_b200")
_spec = importlib.util.spec_from_file_location(
    "crystal_packing_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["crystal_packing_b200"] = _mod
_spec.loader.exec_module(_mod)
