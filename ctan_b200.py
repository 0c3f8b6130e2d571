"""Import alias: `import ctan_b200` loads the package that lives in the (non-identifier)
directory `non-dissipative-propagation-ctdgs_b200/` next to this file."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "non-dissipative-propagation-ctdgs_b200")
_spec = importlib.util.spec_from_file_location(
    "ctan_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ctan_b200"] = _mod
_spec.loader.exec_module(_mod)
