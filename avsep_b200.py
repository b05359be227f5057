"""Import alias: `import avsep_b200` loads the package in
`audio-visual-sepatation-in-visual-agnostic-situtation_b200/` (its directory name is not a valid
Python identifier) and registers it -- and its submodules -- under the short name."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_LONG = "audio-visual-sepatation-in-visual-agnostic-situtation_b200"
_pkg = importlib.import_module(_LONG)
for _name, _mod in list(sys.modules.items()):
    if _name == _LONG or _name.startswith(_LONG + "."):
        sys.modules["avsep_b200" + _name[len(_LONG):]] = _mod
sys.modules[__name__] = _pkg
