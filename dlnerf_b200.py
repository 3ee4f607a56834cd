"""Importable alias of the package directory ``depth-lidar-nerf_b200/`` (hyphens are not valid in an
``import`` statement):  ``import dlnerf_b200 as dn``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("depth-lidar-nerf_b200")
