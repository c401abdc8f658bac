"""Single-layer encoders (reference: multimodn/encoders/slp_encoders.py:5-34).  With no hidden
layer the activation is never applied (mlp_encoder.py:75 iterates an empty list), so all three
compute ``Linear([x || state])``; the classes exist for API parity."""
from typing import Callable

from torch import sigmoid

from .mlp_encoder import MLPEncoder


class SLPEncoder(MLPEncoder):
    def __init__(self, state_size: int, n_features: int, activation: Callable = sigmoid):
        super().__init__(state_size, n_features, (), activation)


def _identity(x):
    return x


class LinearEncoder(SLPEncoder):
    def __init__(self, state_size: int, n_features: int):
        super().__init__(state_size, n_features, _identity)


class LogisticEncoder(SLPEncoder):
    def __init__(self, state_size: int, n_features: int):
        super().__init__(state_size, n_features, sigmoid)
