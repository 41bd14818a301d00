"""Import alias: ``import mi_b200`` loads the package in ``mutual-information-multimodal_b200/``
(the directory name is not a Python identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mutual-information-multimodal_b200")
_spec = importlib.util.spec_from_file_location(
    "mi_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mi_b200"] = _mod
_spec.loader.exec_module(_mod)
