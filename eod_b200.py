"""Import shim: the package directory is ``embodied-object-detection_b200`` (hyphenated); ``import eod_b200``
returns that package object."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("embodied-object-detection_b200")
