"""Import alias: ``import b200seg`` loads the package that lives in the (non-importable, hyphenated)
directory ``cbim-medical-image-segmentation_b200/`` next to this file."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cbim-medical-image-segmentation_b200")
_spec = importlib.util.spec_from_file_location(
    "b200seg", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200seg"] = _mod
_spec.loader.exec_module(_mod)
