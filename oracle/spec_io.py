"""Model-spec helpers shared by the oracle's users.  TEST INFRASTRUCTURE ONLY (see
multimodn_oracle.py header): npz (de)serialisation of a model spec, extraction of a spec from
torch modules that follow the reference's encoder/decoder contracts, and the synthetic
generators for the BASELINE.json configurations.
"""
from __future__ import annotations

import json

import numpy as np


# --------------------------------------------------------------------------------------
# spec <-> flat dict of arrays (for np.savez)
# --------------------------------------------------------------------------------------
def spec_to_arrays(spec, prefix="spec"):
    meta = dict(state_size=int(spec["state_size"]), encoders=[], decoders=[])
    arrays = {f"{prefix}_init": np.asarray(spec["init_state"])}
    for e, enc in enumerate(spec["encoders"]):
        meta["encoders"].append(dict(kind=enc["kind"], n_features=int(enc["n_features"]), act=enc["act"],
                                     dropout=float(enc.get("dropout", 0.0)), n_layers=len(enc["layers"])))
        for j, (W, b) in enumerate(enc["layers"]):
            arrays[f"{prefix}_e{e}_W{j}"] = np.asarray(W)
            arrays[f"{prefix}_e{e}_b{j}"] = np.asarray(b)
    for d, dec in enumerate(spec["decoders"]):
        meta["decoders"].append(dict(n_classes=int(dec["n_classes"]), hidden_act=dec["hidden_act"],
                                     out_act=dec["out_act"], n_layers=len(dec["layers"])))
        for j, (W, b) in enumerate(dec["layers"]):
            arrays[f"{prefix}_d{d}_W{j}"] = np.asarray(W)
            arrays[f"{prefix}_d{d}_b{j}"] = np.asarray(b)
    arrays[f"{prefix}_meta"] = np.array(json.dumps(meta))
    return arrays


def spec_from_arrays(arrays, prefix="spec"):
    meta = json.loads(str(arrays[f"{prefix}_meta"]))
    spec = dict(state_size=meta["state_size"], init_state=np.array(arrays[f"{prefix}_init"]),
                encoders=[], decoders=[])
    for e, m in enumerate(meta["encoders"]):
        layers = [(np.array(arrays[f"{prefix}_e{e}_W{j}"]), np.array(arrays[f"{prefix}_e{e}_b{j}"]))
                  for j in range(m["n_layers"])]
        spec["encoders"].append(dict(kind=m["kind"], n_features=m["n_features"], act=m["act"],
                                     dropout=m["dropout"], layers=layers))
    for d, m in enumerate(meta["decoders"]):
        layers = [(np.array(arrays[f"{prefix}_d{d}_W{j}"]), np.array(arrays[f"{prefix}_d{d}_b{j}"]))
                  for j in range(m["n_layers"])]
        spec["decoders"].append(dict(n_classes=m["n_classes"], hidden_act=m["hidden_act"],
                                     out_act=m["out_act"], layers=layers))
    return spec


def grads_to_arrays(grads, prefix="grad"):
    arrays = {f"{prefix}_init": np.asarray(grads["init_state"])}
    for e, ls in enumerate(grads["encoders"]):
        for j, (W, b) in enumerate(ls):
            arrays[f"{prefix}_e{e}_W{j}"] = np.asarray(W)
            arrays[f"{prefix}_e{e}_b{j}"] = np.asarray(b)
    for d, ls in enumerate(grads["decoders"]):
        for j, (W, b) in enumerate(ls):
            arrays[f"{prefix}_d{d}_W{j}"] = np.asarray(W)
            arrays[f"{prefix}_d{d}_b{j}"] = np.asarray(b)
    return arrays


# --------------------------------------------------------------------------------------
# torch modules (reference's or the product's: same attribute contracts) -> spec
# --------------------------------------------------------------------------------------
def activation_name(fn):
    """Identify an activation callable by probing it (lambdas have no usable name)."""
    import torch
    probe = torch.tensor([-2.0, -0.5, 0.0, 0.75, 3.0])
    with torch.no_grad():
        out = fn(probe)
    for name, ref in (("identity", probe), ("relu", torch.relu(probe)), ("sigmoid", torch.sigmoid(probe)),
                      ("tanh", torch.tanh(probe))):
        if out.shape == ref.shape and torch.allclose(out, ref, rtol=0, atol=1e-7):
            return name
    raise ValueError(f"unsupported activation {fn!r}")


def _lin(layer):
    return (layer.weight.detach().cpu().numpy().copy(), layer.bias.detach().cpu().numpy().copy())


def spec_from_modules(model):
    """model: anything with .init_state.state_value, .encoders, .decoders (MultiModN contract)."""
    import torch.nn as nn
    S = int(model.init_state.state_size)
    spec = dict(state_size=S, init_state=model.init_state.state_value.detach().cpu().numpy().reshape(-1).copy(),
                encoders=[], decoders=[])
    for enc in model.encoders:
        linears = [m for m in enc.layers if isinstance(m, nn.Linear)]
        drops = [m for m in enc.layers if isinstance(m, nn.Dropout)]
        kind = "mimic" if drops else "mlp"
        first_in = linears[0].in_features
        if kind == "mimic" or len(linears) == 1:
            F = first_in - S
        else:
            F = first_in
        spec["encoders"].append(dict(kind=kind, n_features=int(F), act=activation_name(enc.activation),
                                     dropout=float(drops[0].p) if drops else 0.0,
                                     layers=[_lin(l) for l in linears]))
    for dec in model.decoders:
        if hasattr(dec, "fc"):
            spec["decoders"].append(dict(n_classes=int(dec.n_classes), hidden_act="identity",
                                         out_act=activation_name(dec.activation), layers=[_lin(dec.fc)]))
        else:
            spec["decoders"].append(dict(n_classes=int(dec.n_classes),
                                         hidden_act=activation_name(dec.hidden_activation),
                                         out_act=activation_name(dec.output_activation),
                                         layers=[_lin(l) for l in dec.layers]))
    return spec


# --------------------------------------------------------------------------------------
# random specs + synthetic data for the BASELINE.json configs
# --------------------------------------------------------------------------------------
def _linear_init(rng, out_dim, in_dim):
    # nn.Linear default: U(-1/sqrt(in), 1/sqrt(in)) for both weight and bias
    k = 1.0 / np.sqrt(in_dim)
    return (rng.uniform(-k, k, (out_dim, in_dim)).astype(np.float32),
            rng.uniform(-k, k, (out_dim,)).astype(np.float32))


def random_spec(rng, S, features, enc_kind="mimic", enc_hidden=(32, 32), enc_act="relu", dropout=0.0,
                n_decoders=2, dec_hidden=(32, 32), n_classes=2, dec_hidden_act="relu", dec_out_act="sigmoid"):
    spec = dict(state_size=S, init_state=rng.standard_normal(S).astype(np.float32), encoders=[], decoders=[])
    for F in features:
        if enc_kind == "mimic":
            dims = [F + S] + list(enc_hidden) + [S]
            layers = [_linear_init(rng, o, i) for i, o in zip(dims, dims[1:])]
        else:
            dims = [F] + list(enc_hidden) + [S]
            layers = []
            for j, (i, o) in enumerate(zip(dims, dims[1:])):
                layers.append(_linear_init(rng, o, i + (S if j == len(dims) - 2 else 0)))
        spec["encoders"].append(dict(kind=enc_kind, n_features=F, act=enc_act, dropout=dropout, layers=layers))
    for _ in range(n_decoders):
        dims = [S] + list(dec_hidden) + [n_classes]
        spec["decoders"].append(dict(n_classes=n_classes, hidden_act=dec_hidden_act, out_act=dec_out_act,
                                     layers=[_linear_init(rng, o, i) for i, o in zip(dims, dims[1:])]))
    return spec


CONFIGS = {
    # name: (S, features, enc_kind, enc_hidden, n_decoders, dec_hidden, err_penalty, state_change_penalty)
    "c1_titanic": dict(S=1, features=[6], enc_kind="mlp", enc_hidden=(5, 5), n_decoders=1, dec_hidden=(),
                       err_penalty=0.7, state_change_penalty=0.3),
    "c2_mimic": dict(S=64, features=[6, 99, 1024], enc_kind="mimic", enc_hidden=(32, 32), n_decoders=2,
                     dec_hidden=(32, 32), err_penalty=1.0, state_change_penalty=0.3),
    "c3_mnar": dict(S=256, features=[6, 99, 242, 110, 768, 768, 1024, 1024], enc_kind="mimic",
                    enc_hidden=(32, 32), n_decoders=6, dec_hidden=(32, 32), err_penalty=1.0,
                    state_change_penalty=0.3),
    # wide regime (BASELINE.json configs[3]): precision "bf16", 8192 rows per GPU
    "c4_wide": dict(S=1024, features=[1024, 1024, 768, 768], enc_kind="mimic", enc_hidden=(2048, 2048), n_decoders=2,
                    dec_hidden=(2048,), err_penalty=1.0, state_change_penalty=0.3),
}


def config_spec(name, seed=0, dropout=0.0):
    c = CONFIGS[name]
    rng = np.random.default_rng(seed)
    return random_spec(rng, c["S"], c["features"], enc_kind=c["enc_kind"], enc_hidden=c["enc_hidden"],
                       dropout=dropout, n_decoders=c["n_decoders"], dec_hidden=c["dec_hidden"])


def synthetic_batch(rng, features, n_decoders, B, p_pos=0.3, mnar=False, n_classes=2):
    """x_e ~ N(0,1) fp32, y ~ Bernoulli.  mnar=True: whole-modality NaN per row with
    P(miss | y0=1)=0.5, P(miss | y0=0)=0.1 (generalising
    pipelines/mimic/mimic_single_task_mnar_missingness_pipeline.py:142-149)."""
    data = [rng.standard_normal((B, F)).astype(np.float32) for F in features]
    if n_classes == 2:
        y = (rng.random((B, n_decoders)) < (0.5 if mnar else p_pos)).astype(np.int64)
    else:
        y = rng.integers(0, n_classes, (B, n_decoders)).astype(np.int64)
    if mnar:
        p_miss = np.where(y[:, 0] == 1, 0.5, 0.1)
        for x in data:
            miss = rng.random(B) < p_miss
            x[miss, :] = np.nan
    return data, y
