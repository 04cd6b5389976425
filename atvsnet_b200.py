"""Import shim: the product package lives in ``a-tvsnet_b200/`` (hyphenated, as the layout
contract names it); ``import atvsnet_b200`` loads it under an importable name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "a-tvsnet_b200")
_spec = importlib.util.spec_from_file_location("atvsnet_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["atvsnet_b200"] = _mod
_spec.loader.exec_module(_mod)
