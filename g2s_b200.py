"""Import shim: the package directory `gan-2d-to-3d_b200/` is not a valid Python identifier, so this module loads
it under the name `g2s_b200` (`import g2s_b200; g2s_b200.Renderer(...)`)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gan-2d-to-3d_b200")
_spec = importlib.util.spec_from_file_location("g2s_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["g2s_b200"] = _mod
_spec.loader.exec_module(_mod)
