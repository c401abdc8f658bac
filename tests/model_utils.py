"""Build product models (multimodn_b200) from an oracle spec, and tap gradients."""
import numpy as np
import torch
import torch.nn.functional as F

from multimodn_b200 import MultiModN
from multimodn_b200.encoders import MLPEncoder, MIMIC_MLPEncoder
from multimodn_b200.decoders import ClassDecoder, MLPDecoder

ACT = {"relu": F.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh, "identity": lambda x: x}


def model_from_spec(spec, err_penalty, state_change_penalty, device, missing_mode="row", shuffle_mode=False,
                    precision="fp32"):
    S = spec["state_size"]
    encs, decs = [], []
    for e in spec["encoders"]:
        hidden = tuple(int(W.shape[0]) for W, _ in e["layers"][:-1])
        if e["kind"] == "mimic":
            enc = MIMIC_MLPEncoder(S, e["n_features"], hidden, dropout=e.get("dropout", 0.0), activation=ACT[e["act"]])
            lins = [m for m in enc.layers if isinstance(m, torch.nn.Linear)]
        else:
            enc = MLPEncoder(S, e["n_features"], hidden, ACT[e["act"]])
            lins = list(enc.layers)
        for lin, (W, b) in zip(lins, e["layers"]):
            lin.weight.data = torch.from_numpy(np.array(W, dtype=np.float32))
            lin.bias.data = torch.from_numpy(np.array(b, dtype=np.float32))
        encs.append(enc)
    for d in spec["decoders"]:
        if len(d["layers"]) == 1:
            dec = ClassDecoder(S, d["n_classes"], ACT[d["out_act"]])
            lins = [dec.fc]
        else:
            hidden = tuple(int(W.shape[0]) for W, _ in d["layers"][:-1])
            dec = MLPDecoder(S, hidden, d["n_classes"], output_activation=ACT[d["out_act"]],
                             hidden_activation=ACT[d["hidden_act"]])
            lins = list(dec.layers)
        for lin, (W, b) in zip(lins, d["layers"]):
            lin.weight.data = torch.from_numpy(np.array(W, dtype=np.float32))
            lin.bias.data = torch.from_numpy(np.array(b, dtype=np.float32))
        decs.append(dec)
    model = MultiModN(S, encs, decs, err_penalty, state_change_penalty, shuffle_mode=shuffle_mode,
                      device=device, missing_mode=missing_mode, precision=precision)
    model.init_state.state_value.data = torch.from_numpy(
        np.array(spec["init_state"], dtype=np.float32).reshape(1, -1)).to(device)
    return model


def model_spec(model):
    """current weights of a product model as an oracle spec"""
    from oracle.spec_io import spec_from_modules
    return spec_from_modules(model)


class GradTap(torch.optim.Optimizer):
    """Optimizer that leaves the parameters alone and sums the gradients it is handed."""

    def __init__(self, params):
        super().__init__(list(params), {})
        self.acc = {}

    def step(self):
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    self.acc[p] = self.acc.get(p, 0) + p.grad.detach().clone().cpu()


def tapped_flat(model, tap, scale=1.0):
    def g(p):
        v = tap.acc.get(p)
        return (torch.zeros(p.shape) if v is None else v * scale).numpy().ravel()

    parts = [g(model.init_state.state_value)]
    touched = []
    for enc in model.encoders:
        lins = [m for m in enc.layers if isinstance(m, torch.nn.Linear)]
        for l in lins:
            parts += [g(l.weight), g(l.bias)]
        touched.append(lins[0].weight in tap.acc)
    for dec in model.decoders:
        lins = [dec.fc] if hasattr(dec, "fc") else list(dec.layers)
        for l in lins:
            parts += [g(l.weight), g(l.bias)]
    return np.concatenate(parts), np.array(touched)


def batches(data, y, bs, seq=None, device="cpu"):
    out = []
    for i in range(0, y.shape[0], bs):
        item = [[torch.from_numpy(np.ascontiguousarray(x[i:i + bs])).to(device) for x in data],
                torch.from_numpy(y[i:i + bs]).to(device)]
        if seq is not None:
            item.append(torch.from_numpy(np.tile(np.asarray(seq)[None, :], (len(y[i:i + bs]), 1))))
        out.append(tuple(item))
    return out
