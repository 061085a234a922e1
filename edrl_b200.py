"""Import shim: ``import edrl_b200`` loads the package that lives in the (hyphenated, therefore
not directly importable) directory the project layout prescribes."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "robust-multimodal-learning-for-ophthalmic-disease-grading-via-disentangled-representation_b200")
_spec = importlib.util.spec_from_file_location("edrl_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["edrl_b200"] = _mod
_spec.loader.exec_module(_mod)
