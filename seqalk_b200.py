"""Alias for the graft package (its directory name is not a valid Python identifier).

``import seqalk_b200`` == ``importlib.import_module("multiple-sequence-alignment-openmp-openmpi_b200")``.
"""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
PACKAGE_NAME = "multiple-sequence-alignment-openmp-openmpi_b200"
_pkg = importlib.import_module(PACKAGE_NAME)
sys.modules[__name__] = _pkg
