"""B200-native AREAD hot path: multi-field embedding lookup + Hierarchical Expert Integration
under HEMP masks, as hand-written sm_100a CUDA behind the reference's module API.

The directory name follows the repository naming rule and is not a Python identifier; import
it with ``importlib.import_module("aread-multi-domain-recommendation_b200")`` or, with the
repository root on ``sys.path``, through the drop-in modules ``model.aread`` / ``model.layer``.
"""
from . import _lib                      # noqa: F401  (ctypes binding, loads lazily)
from .layer import (BaseModel, CrossNetwork, FeaturesEmbedding, FeaturesLinear,  # noqa: F401
                    MultiLayerPerceptron)
from .aread import AREAD                # noqa: F401

__all__ = ["AREAD", "BaseModel", "CrossNetwork", "FeaturesEmbedding", "FeaturesLinear", "MultiLayerPerceptron"]
