"""Import alias: `import vnl_b200` -> the hyphenated package directory `vnl-brax-imitation_b200/`."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("vnl-brax-imitation_b200")
sys.modules[__name__] = _pkg
