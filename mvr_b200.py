"""Import shim: the package directory is `multi-view-registration_b200/` (not a valid Python identifier),
so `import mvr_b200` loads it under this name; `mvr_b200.synth`, `mvr_b200.build` resolve inside it."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi-view-registration_b200")
_spec = importlib.util.spec_from_file_location("mvr_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mvr_b200"] = _mod
_spec.loader.exec_module(_mod)
