"""Vectorised torch-CPU port of the reference train step.  TEST / BASELINE INFRASTRUCTURE ONLY
(same rules as multimodn_oracle.py: only tests/, smoke() and bench.py's cpu_baseline /
--impl reference legs may import it).

This is the ``ref_vec`` timer of BASELINE.md section 3: the reference's own arithmetic — torch
ops, autograd, torch.optim.Adam, on the host cores — with its three CPython element loops
(`any(x.isnan().flatten())`, `sum(pred == target)`; multimodn/multimodn.py:147,168,183)
replaced by tensor ops and the batch-level skip replaced by the per-row select.  It is what
the reference would cost on a CPU if those loops were vectorised, i.e. a *stronger* baseline than
the reference as shipped (which spends ~75 % of a step in those loops).
``tests/test_oracle_golden.py::test_torch_port_matches_oracle`` pins it to the numpy oracle.

Restated lines: multimodn.py:139-204 (train step), mlp_encoder.py:40-47,74-80, decoders.py:19-20,
42-46, state.py:29-32.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

_ACT = {"identity": lambda z: z, "relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh}


class TorchPort:
    def __init__(self, spec, err_penalty, state_change_penalty_scaled, lr=1e-3):
        self.S = spec["state_size"]
        self.err = err_penalty
        self.scp = state_change_penalty_scaled
        t = lambda a: torch.tensor(np.asarray(a, dtype=np.float32), requires_grad=True)  # noqa: E731
        self.init = t(spec["init_state"])
        self.enc = [dict(kind=e["kind"], act=e["act"], p=float(e.get("dropout", 0.0)),
                         layers=[(t(W), t(b)) for W, b in e["layers"]]) for e in spec["encoders"]]
        self.dec = [dict(hid=d["hidden_act"], out=d["out_act"], C=d["n_classes"],
                         layers=[(t(W), t(b)) for W, b in d["layers"]]) for d in spec["decoders"]]
        self.params = [self.init] + [p for m in self.enc + self.dec for wb in m["layers"] for p in wb]
        self.opt = torch.optim.Adam(self.params, lr)

    def _encoder(self, e, state, x, train):
        m = self.enc[e]
        act = _ACT[m["act"]]
        if m["kind"] == "mimic":
            h = torch.cat([x, state], dim=1)
            if train and m["p"] > 0:
                h = F.dropout(h, m["p"], True)
            for W, b in m["layers"]:
                h = act(F.linear(h, W, b))
            return h
        h = x
        for W, b in m["layers"][:-1]:
            h = act(F.linear(h, W, b))
        W, b = m["layers"][-1]
        return F.linear(torch.cat([h, state], dim=1), W, b)

    def _decoder(self, d, state):
        m = self.dec[d]
        h = state
        for W, b in m["layers"][:-1]:
            h = _ACT[m["hid"]](F.linear(h, W, b))
        W, b = m["layers"][-1]
        return _ACT[m["out"]](F.linear(h, W, b))

    def forward_loss(self, data, y, train=True):
        """returns (loss, ce (E+1,D) tensor, state_change (E,), n_correct)"""
        B = y.shape[0]
        E, D = len(self.enc), len(self.dec)
        state = self.init.unsqueeze(0).expand(B, -1)
        ce = torch.zeros(E + 1, D)
        n_correct = torch.zeros(E + 1, D)
        sc = torch.zeros(E)

        def heads(row, mask):
            for d in range(D):
                out = self._decoder(d, state)
                per_row = F.cross_entropy(out, y[:, d], reduction="none")
                ce[row, d] = (per_row * mask).sum() / B
                n_correct[row, d] = ((out.argmax(dim=1) == y[:, d]) & (mask > 0)).sum()

        heads(0, torch.ones(B))
        for e in range(E):
            x = data[e]
            nan = torch.isnan(x)
            present = ~nan.any(dim=1)
            new = self._encoder(e, state, torch.where(nan, torch.zeros_like(x), x), train)
            new_state = torch.where(present[:, None], new, state)
            sc[e] = ((new_state - state) ** 2).mean()
            state = new_state
            heads(e + 1, present.float())
        loss = ce.sum() / (D * (E + 1)) * self.err + sc.sum() / E * self.scp
        return loss, ce, sc, n_correct

    def train_step(self, data, y):
        self.opt.zero_grad()
        loss, ce, sc, _ = self.forward_loss(data, y, True)
        loss.backward()
        self.opt.step()
        return float(loss.detach())

    def grads(self, data, y, train=False):
        self.opt.zero_grad()
        loss, ce, sc, _ = self.forward_loss(data, y, train)
        loss.backward()
        return float(loss.detach()), ce.detach().numpy(), sc.detach().numpy(), \
            [None if p.grad is None else p.grad.numpy().copy() for p in self.params]
