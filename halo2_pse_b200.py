"""Import shim: exposes the package directory `halo2-pse_b200/` (hyphenated by
the repo's naming contract) under the importable name `halo2_pse_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "halo2-pse_b200")
_spec = importlib.util.spec_from_file_location(
    "halo2_pse_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["halo2_pse_b200"] = _mod
_spec.loader.exec_module(_mod)
