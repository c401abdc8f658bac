"""Encoder contract (reference: multimodn/encoders/multimod_encoder.py:8-17)."""
from abc import ABC, abstractmethod

from torch import Tensor, nn


class MultiModEncoder(nn.Module, ABC):
    """``forward(state, x) -> new state``; carries ``state_size``."""

    def __init__(self, state_size: int):
        super().__init__()
        self.state_size = state_size

    @abstractmethod
    def forward(self, state: Tensor, x: Tensor) -> Tensor:
        ...
