"""Import alias: the package directory is ``low-cost-mocap_b200`` (not a Python identifier);
``import mocap_b200`` gives the same module object."""
import importlib
import sys

_pkg = importlib.import_module("low-cost-mocap_b200")
sys.modules[__name__] = _pkg
