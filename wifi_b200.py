"""Import shim: the package directory is named after the reference repository and contains a
hyphen, so `import wifi_b200` resolves it through importlib."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("gnuradio-wifi-imagetransfer_b200")
sys.modules[__name__] = _pkg
