"""Import alias: the package directory is named `pretrain-gnns_b200` (not a Python identifier)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("pretrain-gnns_b200")
