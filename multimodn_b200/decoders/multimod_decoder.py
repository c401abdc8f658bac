"""Decoder contract (reference: multimodn/decoders/multimod_decoder.py:7-16)."""
from abc import ABC, abstractmethod

from torch import Tensor, nn


class MultiModDecoder(nn.Module, ABC):
    """``forward(state) -> (B, n_classes)``; carries ``state_size`` and ``n_classes``."""

    def __init__(self, state_size: int):
        super().__init__()
        self.state_size = state_size

    @abstractmethod
    def forward(self, state: Tensor) -> Tensor:
        ...
