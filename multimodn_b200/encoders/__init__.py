from .multimod_encoder import MultiModEncoder
from .mlp_encoder import MLPEncoder, MIMIC_MLPEncoder, MLPFeatureEncoder
from .slp_encoders import SLPEncoder, LinearEncoder, LogisticEncoder

__all__ = ["MultiModEncoder", "MLPEncoder", "MIMIC_MLPEncoder", "MLPFeatureEncoder", "SLPEncoder",
           "LinearEncoder", "LogisticEncoder"]
