"""`import vatlq` -> the package in ./vatl4pose-wacv2024_b200 (whose directory name is not an identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_REAL = "vatl4pose-wacv2024_b200"
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name == _REAL or _name.startswith(_REAL + "."):
        sys.modules["vatlq" + _name[len(_REAL):]] = _mod
sys.modules[__name__] = _pkg
