"""Initial state of the fusion chain (reference: multimodn/state.py).

``TrainableInitState`` keeps the reference's parameter name and shape
(``state_value`` of shape ``(1, state_size)``, state.py:25-27) so reference checkpoints load.
Inside the fused step the state is never tiled to ``(B, S)``: every batch tile broadcasts it
from the packed parameter buffer, and its gradient is the column sum of dLoss/ds_0.
"""
from abc import ABC, abstractmethod
from itertools import cycle
from typing import List, Optional

import torch
from torch import Tensor, nn


class InitState(nn.Module, ABC):
    def __init__(self, state_size: int):
        super().__init__()
        self.state_size = state_size

    @abstractmethod
    def forward(self, batch_size) -> Tensor:
        ...


class TrainableInitState(InitState):
    """Learned (1, S) row broadcast over the batch (state.py:19-32)."""

    def __init__(self, state_size: int, device: Optional[torch.device] = None):
        super().__init__(state_size)
        self.device = device
        self.state_value = nn.Parameter(torch.randn((1, state_size), device=device))

    def forward(self, batch_size) -> Tensor:
        return self.state_value.expand(batch_size, -1).clone()


class StaticInitState(InitState):
    """Fixed per-sample states cycled from a list (state.py:34-47).  Kept for API completeness;
    the fused step supports the trainable state only (no reference pipeline uses this one)."""

    def __init__(self, states: List[Tensor]):
        super().__init__(states[0].size(0))
        self._states = cycle(states)

    def forward(self, batch_size) -> Tensor:
        rows = [next(self._states).reshape(1, -1) for _ in range(batch_size)]
        return torch.cat(rows, dim=0).detach()
