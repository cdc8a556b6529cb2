"""Drop-in for the reference `model/layer.py` (namespace-package shadowing).

Put this repository's root ahead of the reference tree on `sys.path`: `model.layer` and
`model.aread` then resolve here while every other `model.*`, `run`, `config`, ... still resolves
to the reference.  Do NOT add a `model/__init__.py`.
"""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

_impl = importlib.import_module("aread-multi-domain-recommendation_b200.layer")

BaseModel = _impl.BaseModel
FeaturesEmbedding = _impl.FeaturesEmbedding
FeaturesLinear = _impl.FeaturesLinear
MultiLayerPerceptron = _impl.MultiLayerPerceptron
CrossNetwork = _impl.CrossNetwork


def __getattr__(name):          # baseline-only layers are served from the reference tree
    return getattr(_impl, name)
