"""multimodn_b200 — the sequential-fusion step of EPFLiGHT/MultiModN on NVIDIA B200 (sm_100a).

Same import surface as the reference's ``multimodn`` package (``multimodn/__init__.py:1-3``):

    from multimodn_b200 import MultiModN, MultiModNHistory
    from multimodn_b200.encoders import MLPEncoder, MIMIC_MLPEncoder
    from multimodn_b200.decoders import LogisticDecoder, MLPDecoder
"""
from .history import MultiModNHistory, display_title
from .multimodn import MultiModN, get_performance_metrics, performance_metrics
from .state import InitState, TrainableInitState, StaticInitState
from .optim import FusedAdam

__all__ = ["MultiModN", "MultiModNHistory", "display_title", "get_performance_metrics", "performance_metrics",
           "InitState", "TrainableInitState", "StaticInitState", "FusedAdam"]
