"""tcvom_b200 -- B200-native (sm_100a) implementation of TCVOM's GCA+TAM frame-window hot path.

Public surface mirrors the reference's operator/plugin interface for this path:
``get_VMN_models``, ``VMN``, ``FeatureAggregationModule``, ``GuidedCxtAtten``, ``EvalModel``.
"""
from .model import EvalModel, FeatureAggregationModule, GuidedCxtAtten, VMN, get_VMN_models  # noqa: F401

__version__ = "0.1.0"
