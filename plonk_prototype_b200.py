"""Import shim: the package directory is `plonk-prototype_b200/` (a hyphen is not importable), so this
module loads it under the name `plonk_prototype_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "plonk-prototype_b200")
_spec = importlib.util.spec_from_file_location("plonk_prototype_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["plonk_prototype_b200"] = _mod
_spec.loader.exec_module(_mod)
