"""Dense encoders (reference: multimodn/encoders/mlp_encoder.py).

Parameter names and shapes match the reference's ``state_dict`` (``layers.{j}.weight/bias``;
in ``MIMIC_MLPEncoder`` ``layers.0`` is the Dropout, so the Linears are ``layers.1..``).
``forward`` is the standalone module contract; ``MultiModN`` does not call it — it lowers the
module to a layer plan (``multimodn_b200/plan.py``) executed by the fused CUDA step.
"""
from typing import Callable, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .multimod_encoder import MultiModEncoder


class MLPEncoder(MultiModEncoder):
    """Hidden layers act on the features only; the state is concatenated (after the hidden
    representation) to the input of the LAST layer, which has no activation
    (mlp_encoder.py:61-80)."""

    _mmn_kind = "mlp"

    def __init__(self, state_size: int, n_features: int, hidden_layers: Tuple[int, ...],
                 activation: Callable = F.relu, device: Optional[torch.device] = None):
        super().__init__(state_size)
        self.activation = activation
        widths = [n_features, *hidden_layers, state_size]
        self.layers = nn.ModuleList()
        n_lin = len(widths) - 1
        for j in range(n_lin):
            fan_in = widths[j] + (state_size if j == n_lin - 1 else 0)
            self.layers.append(nn.Linear(fan_in, widths[j + 1], device=device))

    def forward(self, state: Tensor, x: Tensor) -> Tensor:
        *hidden, head = self.layers
        for lin in hidden:
            x = self.activation(lin(x))
        return head(torch.cat((x, state), dim=1))


class MIMIC_MLPEncoder(MultiModEncoder):
    """[x || state] -> Dropout -> (Linear -> activation) for EVERY layer, the last included
    (mlp_encoder.py:27-47)."""

    _mmn_kind = "mimic"

    def __init__(self, state_size: int, n_features: int, hidden_layers: Tuple[int, ...], dropout: float = .2,
                 activation: Callable = F.relu, device: Optional[torch.device] = None):
        super().__init__(state_size)
        self.activation = activation
        self.dropout = dropout
        widths = [n_features + state_size, *hidden_layers, state_size]
        self.layers = nn.ModuleList([nn.Dropout(dropout)])
        for fan_in, fan_out in zip(widths, widths[1:]):
            self.layers.append(nn.Linear(fan_in, fan_out, device=device))

    def forward(self, state: Tensor, x: Tensor) -> Tensor:
        h = self.layers[0](torch.cat((x, state), dim=1))
        for lin in list(self.layers)[1:]:
            h = self.activation(lin(h))
        return h


class MLPFeatureEncoder(MLPEncoder):
    """One scalar feature through one hidden layer (mlp_encoder.py:81-94)."""

    def __init__(self, state_size: int, hidden_size: int, activation: Callable = F.relu,
                 device: Optional[torch.device] = None):
        super().__init__(state_size, 1, (hidden_size,), activation, device)

    def forward(self, state: Tensor, x) -> Tensor:
        return super().forward(state, torch.as_tensor(x, dtype=state.dtype, device=state.device))
