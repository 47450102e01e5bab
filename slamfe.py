"""Import alias: `import slamfe` loads the package directory `67604-slam---video-navigation_b200/`
(whose name is not a valid Python identifier) under the module name `slamfe`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "67604-slam---video-navigation_b200")
_spec = importlib.util.spec_from_file_location(
    "slamfe", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["slamfe"] = _mod
_spec.loader.exec_module(_mod)
