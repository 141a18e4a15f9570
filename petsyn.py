"""Import shim: the product package lives in the directory
``causality-informed-pet-synthesis-from-multi-modal-data_b200/`` whose name is not a Python identifier.
``import petsyn`` loads that directory as the package ``petsyn_b200`` and re-exports it."""
import importlib.util
import os
import sys

PACKAGE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                           "causality-informed-pet-synthesis-from-multi-modal-data_b200")
_NAME = "petsyn_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(PACKAGE_DIR, "__init__.py"), submodule_search_locations=[PACKAGE_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

_pkg = sys.modules[_NAME]
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
