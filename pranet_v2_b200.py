"""Import shim: the package directory is named `pranet-v2_b200/` (not an importable identifier), so
`import pranet_v2_b200` resolves here and this module replaces itself with that package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pranet-v2_b200")
_spec = importlib.util.spec_from_file_location(
    "pranet_v2_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pranet_v2_b200"] = _mod
_spec.loader.exec_module(_mod)
