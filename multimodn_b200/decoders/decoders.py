"""Decoder heads (reference: multimodn/decoders/decoders.py).  ``state_dict`` keys match the
reference (``fc.*`` for ClassDecoder, ``layers.{j}.*`` for MLPDecoder)."""
from typing import Callable, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn, sigmoid

from .multimod_decoder import MultiModDecoder


class ClassDecoder(MultiModDecoder):
    """``activation(Linear(state))`` (decoders.py:9-20)."""

    def __init__(self, state_size: int, n_classes: int, activation: Callable,
                 device: Optional[torch.device] = None):
        super().__init__(state_size)
        self.n_classes = n_classes
        self.fc = nn.Linear(state_size, n_classes, device=device)
        self.activation = activation

    def forward(self, state: Tensor) -> Tensor:
        return self.activation(self.fc(state))


class MLPDecoder(MultiModDecoder):
    """ReLU MLP with a squashed output layer (decoders.py:22-46)."""

    def __init__(self, state_size: int, hidden_layers: Tuple[int, ...], n_classes: int = 2,
                 output_activation: Callable = sigmoid, hidden_activation: Callable = F.relu,
                 device: Optional[torch.device] = None):
        super().__init__(state_size)
        self.output_activation = output_activation
        self.hidden_activation = hidden_activation
        self.n_classes = n_classes
        widths = [state_size, *hidden_layers, n_classes]
        self.layers = nn.ModuleList(nn.Linear(i, o, device=device) for i, o in zip(widths, widths[1:]))

    def forward(self, x: Tensor) -> Tensor:
        *hidden, head = self.layers
        for lin in hidden:
            x = self.hidden_activation(lin(x))
        return self.output_activation(head(x))


class LogisticDecoder(ClassDecoder):
    """Two sigmoid outputs (decoders.py:49-53)."""

    def __init__(self, state_size: int, device: Optional[torch.device] = None):
        super().__init__(state_size, 2, sigmoid, device)
