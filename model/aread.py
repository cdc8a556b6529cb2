"""Drop-in for the reference `model/aread.py` (see model/layer.py in this directory)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

AREAD = importlib.import_module("aread-multi-domain-recommendation_b200.aread").AREAD
