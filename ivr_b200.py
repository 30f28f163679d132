"""Import alias: ``import ivr_b200`` -> the package in
``intelligent-video-analysis-retrieval-system_b200/`` (a hyphenated directory
name cannot be written in an ``import`` statement)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("intelligent-video-analysis-retrieval-system_b200")
sys.modules[__name__] = _pkg
for _name, _mod in list(sys.modules.items()):
    if _name.startswith("intelligent-video-analysis-retrieval-system_b200."):
        sys.modules["ivr_b200." + _name.split(".", 1)[1]] = _mod
