"""Import shim: ``import asr_b200`` loads the package that lives in ``asr-using-robust-nn_b200/``.

The package directory carries the repository's name (with hyphens, so Python cannot import it by
that name); this module registers it under the importable name ``asr_b200``.
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "asr-using-robust-nn_b200")
_spec = importlib.util.spec_from_file_location("asr_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["asr_b200"] = _mod
_spec.loader.exec_module(_mod)
