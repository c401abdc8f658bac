"""Lowering of encoder / decoder modules to the layer plan executed by libmmn.so, and packing of
their parameters into one contiguous buffer.

The reference walks ``nn.ModuleList``s and calls each module's ``forward``
(multimodn/multimodn.py:141-143, 159-163, 173, 176-178).  Here the modules are inspected once:
every supported module is a stack of ``nn.Linear`` layers with a known activation and a known
place where the running state is concatenated, which is what ``mmn_layer_desc`` (include/mmn.h)
describes.  Anything else — recurrent or convolutional encoders, unknown activations — raises:
there is no eager fallback.

Parameters stay ordinary ``nn.Parameter``s under the reference's ``state_dict`` names; their
storage is re-pointed at slices of one flat fp32 buffer so the kernels (and the fused Adam) see
a single array while ``optimizer.step()``, ``state_dict()`` and ``load_state_dict()`` keep working.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List

import torch
from torch import nn

from . import _lib


def activation_name(fn) -> str:
    """Identify an activation callable by probing it (lambdas have no usable name)."""
    probe = torch.tensor([-2.0, -0.5, 0.0, 0.75, 3.0])
    try:
        with torch.no_grad():
            out = fn(probe)
    except Exception as exc:  # noqa: BLE001
        raise NotImplementedError(f"activation {fn!r} could not be probed: {exc}") from exc
    table = {"identity": probe, "relu": torch.relu(probe), "sigmoid": torch.sigmoid(probe),
             "tanh": torch.tanh(probe)}
    for name, ref in table.items():
        if torch.is_tensor(out) and out.shape == ref.shape and torch.allclose(out, ref, rtol=0, atol=1e-7):
            return name
    raise NotImplementedError(f"unsupported activation {fn!r}: the fused step implements {sorted(table)}")


@dataclass
class LoweredLayer:
    linear: nn.Linear
    in_dim: int
    out_dim: int
    act: str
    has_state: bool
    w_off: int = 0
    b_off: int = 0


@dataclass
class LoweredModule:
    kind: str                     # 'mlp' | 'mimic' | 'decoder'
    layers: List[LoweredLayer] = field(default_factory=list)
    n_features: int = 0
    dropout: float = 0.0
    n_classes: int = 0


def _kind_of(module) -> str:
    kind = getattr(module, "_mmn_kind", None)
    if kind:
        return kind
    names = {c.__name__ for c in type(module).__mro__}        # the reference's own classes
    if "MIMIC_MLPEncoder" in names:
        return "mimic"
    if "MLPEncoder" in names:
        return "mlp"
    raise NotImplementedError(
        f"{type(module).__name__} is not supported by the fused step: dense encoders only "
        "(MLPEncoder, MIMIC_MLPEncoder, MLPFeatureEncoder, SLPEncoder, LinearEncoder, LogisticEncoder). "
        "Recurrent and convolutional encoders are out of scope (their rows are not independent).")


def lower_encoder(enc, S: int) -> LoweredModule:
    kind = _kind_of(enc)
    linears = [m for m in enc.layers if isinstance(m, nn.Linear)]
    others = [m for m in enc.layers if not isinstance(m, (nn.Linear, nn.Dropout))]
    if others or not linears:
        raise NotImplementedError(f"{type(enc).__name__}: unsupported layer {type(others[0]).__name__ if others else None}")
    act = activation_name(enc.activation)
    out = LoweredModule(kind=kind)
    n = len(linears)
    for j, lin in enumerate(linears):
        if kind == "mimic":           # mlp_encoder.py:27-47
            has_state, a = j == 0, act
        else:                         # mlp_encoder.py:61-80
            has_state, a = j == n - 1, (act if j < n - 1 else "identity")
        out.layers.append(LoweredLayer(lin, lin.in_features - (S if has_state else 0), lin.out_features, a, has_state))
    out.n_features = out.layers[0].in_dim
    if kind == "mimic":
        drops = [m for m in enc.layers if isinstance(m, nn.Dropout)]
        out.dropout = float(drops[0].p) if drops else 0.0
    if out.layers[-1].out_dim != S:
        raise ValueError(f"{type(enc).__name__}: last layer produces {out.layers[-1].out_dim}, state size is {S}")
    return out


def lower_decoder(dec, S: int) -> LoweredModule:
    if hasattr(dec, "fc") and isinstance(dec.fc, nn.Linear):          # decoders.py:9-20
        layers = [LoweredLayer(dec.fc, dec.fc.in_features, dec.fc.out_features, activation_name(dec.activation), False)]
    elif hasattr(dec, "layers") and all(isinstance(m, nn.Linear) for m in dec.layers):   # decoders.py:22-46
        hid, outa = activation_name(dec.hidden_activation), activation_name(dec.output_activation)
        linears = list(dec.layers)
        layers = [LoweredLayer(l, l.in_features, l.out_features, hid if j < len(linears) - 1 else outa, False)
                  for j, l in enumerate(linears)]
    else:
        raise NotImplementedError(f"{type(dec).__name__} is not supported by the fused step "
                                  "(ClassDecoder, LogisticDecoder, MLPDecoder)")
    if layers[0].in_dim != S:
        raise ValueError(f"{type(dec).__name__}: expects a state of {layers[0].in_dim}, state size is {S}")
    n_classes = int(dec.n_classes)
    if layers[-1].out_dim != n_classes:
        raise ValueError(f"{type(dec).__name__}: last layer width {layers[-1].out_dim} != n_classes {n_classes}")
    return LoweredModule(kind="decoder", layers=layers, n_classes=n_classes)


def _align4(n: int) -> int:
    return (n + 3) & ~3


class PackedModel:
    """Offsets of every parameter in the flat buffer + the ctypes model description."""

    def __init__(self, init_param: nn.Parameter, encoders, decoders, S: int):
        if len(encoders) > _lib.MAX_ENCODERS or len(decoders) > _lib.MAX_DECODERS:
            raise NotImplementedError(f"at most {_lib.MAX_ENCODERS} encoders and {_lib.MAX_DECODERS} decoders")
        self.S = S
        self.init_param = init_param
        self.encoders = [lower_encoder(e, S) for e in encoders]
        self.decoders = [lower_decoder(d, S) for d in decoders]
        off = 0
        self.slots = []                     # (parameter, offset, owner) ; owner: -1 shared, e encoder id
        self.init_off = off
        self.slots.append((init_param, off, -1))
        off = _align4(off + S)
        for e, m in enumerate(self.encoders):
            for l in m.layers:
                l.w_off = off
                self.slots.append((l.linear.weight, off, e))
                off = _align4(off + l.linear.weight.numel())
                l.b_off = off
                self.slots.append((l.linear.bias, off, e))
                off = _align4(off + l.linear.bias.numel())
        for m in self.decoders:
            for l in m.layers:
                l.w_off = off
                self.slots.append((l.linear.weight, off, -1))
                off = _align4(off + l.linear.weight.numel())
                l.b_off = off
                self.slots.append((l.linear.bias, off, -1))
                off = _align4(off + l.linear.bias.numel())
        self.n_params = off
        seen = set()
        for p, _, _ in self.slots:
            if id(p) in seen:
                raise NotImplementedError("parameters shared between layers are not supported")
            seen.add(id(p))

    def model_desc(self, precision: int = 0):
        """-> (ModelDesc, keep-alive tuple); precision: 0 = fp32 (fused step kernels), 1 = bf16 (wide regime)"""
        E, D = len(self.encoders), len(self.decoders)
        encs = (_lib.EncoderDesc * E)()
        decs = (_lib.DecoderDesc * D)()

        def fill(dst, layers):
            if len(layers) > _lib.MAX_LAYERS:
                raise NotImplementedError(f"at most {_lib.MAX_LAYERS} Linear layers per module")
            for j, l in enumerate(layers):
                dst[j].in_dim, dst[j].out_dim = l.in_dim, l.out_dim
                dst[j].act, dst[j].has_state = _lib.ACT_CODES[l.act], int(l.has_state)
                dst[j].w_off, dst[j].b_off = l.w_off, l.b_off

        for e, m in enumerate(self.encoders):
            encs[e].n_features, encs[e].n_layers, encs[e].dropout_p = m.n_features, len(m.layers), m.dropout
            fill(encs[e].layers, m.layers)
        for d, m in enumerate(self.decoders):
            if m.n_classes > _lib.MAX_CLASSES:
                raise NotImplementedError(f"at most {_lib.MAX_CLASSES} classes per decoder")
            decs[d].n_classes, decs[d].n_layers = m.n_classes, len(m.layers)
            fill(decs[d].layers, m.layers)
        desc = _lib.ModelDesc(self.S, E, D, int(precision), self.init_off, self.n_params,
                              C.cast(encs, C.POINTER(_lib.EncoderDesc)), C.cast(decs, C.POINTER(_lib.DecoderDesc)))
        return desc, (encs, decs)

    def encoder_range(self, e: int):
        """[lo, hi) of encoder e's parameters in the packed buffer (contiguous: layers are packed in order)"""
        offs = [(off, off + p.numel()) for p, off, owner in self.slots if owner == e]
        return min(lo for lo, _ in offs), max(hi for _, hi in offs)

    def encoder_layer_ranges(self, e: int):
        """[lo, hi) of every Linear layer (weight and bias, adjacent in the packed buffer) of encoder e, in layer order"""
        return [(l.w_off, l.b_off + l.linear.bias.numel()) for l in self.encoders[e].layers]

    def decoder_range(self):
        """[lo, hi) of all decoders' parameters (packed behind the last encoder)"""
        offs = [(l.w_off, l.b_off + l.linear.bias.numel()) for m in self.decoders for l in m.layers]
        return min(lo for lo, _ in offs), max(hi for _, hi in offs)

    def complement_ranges(self, encoders, total: int, also=()):
        """what is left of [0, total) once the blocks of `encoders` (and the ranges in `also`) are removed, as a list of [lo, hi)"""
        taken = sorted([self.encoder_range(e) for e in encoders] + list(also))
        out, at = [], 0
        for lo, hi in taken:
            if lo > at:
                out.append((at, lo))
            at = max(at, hi)
        if at < total:
            out.append((at, total))
        return out

    # -- flat buffer management ----------------------------------------------------------------
    def pack(self, device) -> torch.Tensor:
        flat = torch.zeros(self.n_params, dtype=torch.float32, device=device)
        with torch.no_grad():
            for p, off, _ in self.slots:
                view = flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.detach().to(device=device, dtype=torch.float32))
                p.data = view
        return flat

    def is_packed(self, flat: torch.Tensor) -> bool:
        base = flat.data_ptr()
        return all(p.data_ptr() == base + 4 * off and p.dtype == torch.float32 for p, off, _ in self.slots)
