"""CPU oracle for the MultiModN sequential-fusion step.  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference algorithm.  It exists to CHECK the
CUDA path; it is never the thing that is shipped or measured.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import it.  The product package ``multimodn_b200`` never imports anything from
``oracle/`` and raises if its CUDA library is missing.

Parity status: PINNED.  The reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the unmodified reference
executed in the build container: ``tests/golden/make_golden.py`` imports
``/root/reference/multimodn`` and dumps inputs, weights, per-step losses, gradients,
history matrices, predictions and states into ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays every fixture through this file.

Reference lines restated (paths relative to /root/reference):

* sequence resolution ........ multimodn/multimodn.py:509-531      -> resolve_sequence
* train step ................. multimodn/multimodn.py:139-204      -> forward / train_step
* test step .................. multimodn/multimodn.py:301-357      -> forward(train=False)
* predict .................... multimodn/multimodn.py:434-455      -> predict
* get_states ................. multimodn/multimodn.py:476-492      -> get_states
* epoch bookkeeping .......... multimodn/multimodn.py:104-115,206-250,269-280,360-409 -> EpochAccumulator
* confusion cells ............ multimodn/multimodn.py:51-63        -> forward (tp/tn/fp/fn)
* initial state .............. multimodn/state.py:29-32            -> forward (s0 tile)
* MLPEncoder ................. multimodn/encoders/mlp_encoder.py:49-80   -> expand_encoder('mlp')
* MIMIC_MLPEncoder ........... multimodn/encoders/mlp_encoder.py:9-47    -> expand_encoder('mimic')
* SLP/Linear/Logistic ........ multimodn/encoders/slp_encoders.py:5-34   -> 'mlp' with no hidden layer
* ClassDecoder / MLPDecoder .. multimodn/decoders/decoders.py:9-53       -> decoder_forward
* criterion .................. torch.nn.CrossEntropyLoss() default, applied to the decoder
                               OUTPUTS (already squashed), as every pipeline does
                               (pipelines/titanic/titanic_mlp_pipeline.py:76)

Missingness: ``missing_mode='batch'`` is the reference rule verbatim (any NaN in a modality
tensor skips that encoder for the whole batch, multimodn.py:167-169).  ``missing_mode='row'``
is the north-star rule (per-row select); it equals the reference evaluated one row at a time
(batch size 1) and averaged, which is how the golden fixtures pin it.

The backward pass is written out by hand (no autograd) in the same order the CUDA kernel
uses, so that a mismatch can be localised layer by layer.
"""
from __future__ import annotations

import numpy as np

ACTS = ("identity", "relu", "sigmoid", "tanh")


# --------------------------------------------------------------------------------------
# activations
# --------------------------------------------------------------------------------------
def act_fwd(name, z):
    if name == "identity":
        return z
    if name == "relu":
        return np.maximum(z, 0)
    if name == "sigmoid":
        return (1.0 / (1.0 + np.exp(-z))).astype(z.dtype)
    if name == "tanh":
        return np.tanh(z)
    raise ValueError(f"unsupported activation {name!r}")


def act_bwd(name, out, dout):
    """derivative expressed through the activation OUTPUT (what the kernel stashes)."""
    if name == "identity":
        return dout
    if name == "relu":
        return dout * (out > 0)
    if name == "sigmoid":
        return dout * out * (1 - out)
    if name == "tanh":
        return dout * (1 - out * out)
    raise ValueError(name)


# --------------------------------------------------------------------------------------
# dropout mask: counter-based hash shared bit-for-bit with the CUDA kernel
# (multimodn_b200/csrc/mmn_kernels.cuh: mmn_dropout_keep)
# --------------------------------------------------------------------------------------
def dropout_keep(seed, enc_id, rows, n_cols, p):
    """keep-mask (bool, (len(rows), n_cols)) for MIMIC_MLPEncoder's dropout on [x || state].

    Restates nn.Dropout(p) at mlp_encoder.py:33-34,43-44 up to the choice of random stream
    (torch's Philox stream cannot be reproduced outside torch; parity is checked with this
    explicit mask on both sides).
    """
    rows = np.asarray(rows, dtype=np.uint32)[:, None]
    cols = np.arange(n_cols, dtype=np.uint32)[None, :]
    with np.errstate(over="ignore"):
        h = np.uint32(seed) ^ (np.uint32(enc_id) * np.uint32(0x9E3779B9))
        x = rows * np.uint32(0x85EBCA6B) + (cols >> np.uint32(1)) * np.uint32(0xC2B2AE35) + h
        x ^= x >> np.uint32(16)
        x *= np.uint32(0x7FEB352D)
        x ^= x >> np.uint32(15)
        x *= np.uint32(0x846CA68B)
        x ^= x >> np.uint32(16)
    # one 32-bit hash serves a pair of adjacent columns: low 16 bits -> even column, high -> odd
    half = np.where((cols & np.uint32(1)) == 1, x >> np.uint32(16), x & np.uint32(0xFFFF))
    thr = np.uint32(int(np.float32(p) * np.float32(65536.0)))
    return half >= thr


# --------------------------------------------------------------------------------------
# model spec helpers
# --------------------------------------------------------------------------------------
def expand_encoder(enc, S):
    """unified layer list: each layer = dict(W, b, in_dim, out_dim, act, has_state)."""
    kind = enc["kind"]
    layers = []
    n = len(enc["layers"])
    for j, (W, b) in enumerate(enc["layers"]):
        if kind == "mimic":      # mlp_encoder.py:27-38,40-47: [x||state] first, act after every layer
            has_state = j == 0
            act = enc["act"]
        elif kind == "mlp":      # mlp_encoder.py:61-80: state joins the LAST layer, no act on it
            has_state = j == n - 1
            act = enc["act"] if j < n - 1 else "identity"
        else:
            raise ValueError(kind)
        in_dim = W.shape[1] - (S if has_state else 0)
        layers.append(dict(W=W, b=b, in_dim=in_dim, out_dim=W.shape[0], act=act, has_state=has_state))
    assert layers[0]["in_dim"] == enc["n_features"], (layers[0]["in_dim"], enc["n_features"])
    assert layers[-1]["out_dim"] == S
    return layers


def expand_decoder(dec):
    n = len(dec["layers"])
    layers = []
    for j, (W, b) in enumerate(dec["layers"]):
        act = dec["out_act"] if j == n - 1 else dec["hidden_act"]
        layers.append(dict(W=W, b=b, in_dim=W.shape[1], out_dim=W.shape[0], act=act, has_state=False))
    assert layers[-1]["out_dim"] == dec["n_classes"]
    return layers


def cast_spec(spec, dtype):
    out = dict(state_size=spec["state_size"], init_state=spec["init_state"].astype(dtype))
    if "precision" in spec:
        out["precision"] = spec["precision"]
    out["encoders"] = [dict(e, layers=[(W.astype(dtype), b.astype(dtype)) for W, b in e["layers"]])
                       for e in spec["encoders"]]
    out["decoders"] = [dict(d, layers=[(W.astype(dtype), b.astype(dtype)) for W, b in d["layers"]])
                       for d in spec["decoders"]]
    return out


def resolve_sequence(encoder_sequence, n_encoders):
    """multimodn.py:509-531 -> list of (data_idx, enc_idx).  Accepts None, 1-D, or (B, L)."""
    if encoder_sequence is None:
        return [(i, i) for i in range(n_encoders)]
    seq = np.asarray(encoder_sequence)
    if seq.ndim == 2:
        if not (seq == seq[0]).all():
            raise ValueError("Encoder sequence has different values across the batch. "
                             "Hint: set batch size to 1 to avoid this error.")
        seq = seq[0]
    return [(i, int(e)) for i, e in enumerate(seq)]


# --------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------
def round_bf16(x):
    """round-to-nearest-even to bfloat16, returned in the input's float32 container"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    r = ((u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)).view(np.float32)
    return np.where(np.isfinite(x), r, x)


def _rounding(spec):
    """spec['precision'] == 'bf16' restates the wide regime (multimodn_b200/csrc/mmn_wide_step.cuh): the same
    algorithm with weights, layer inputs, activations, states and layer gradients rounded to bfloat16 where the
    CUDA path stores them as bfloat16; sums, the state gradient and the parameter gradients stay float32."""
    return round_bf16 if spec.get("precision") == "bf16" else None


def _mlp_forward(layers, a, s_prev, keep_scale, rnd=None, round_last=True):
    cache = []
    for j, L in enumerate(layers):
        inp = np.concatenate([a, s_prev], axis=1) if L["has_state"] else a
        if j == 0 and keep_scale is not None:
            inp = inp * keep_scale
        W = L["W"]
        if rnd is not None:
            inp, W = rnd(inp), rnd(W)
        z = inp @ W.T + L["b"]
        a = act_fwd(L["act"], z)
        if rnd is not None and (round_last or j < len(layers) - 1):
            a = rnd(a)
        cache.append((inp, a))
    return a, cache


def _ce_rows(p, y):
    """per-row CrossEntropyLoss term on outputs p: logsumexp(p) - p[y]."""
    m = p.max(axis=1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(p - m).sum(axis=1))
    return lse - p[np.arange(p.shape[0]), y]


def _first_argmax(p):
    # torch.max(dim=1) returns the first maximal index (multimodn.py:144); np.argmax too.
    return np.argmax(p, axis=1)


def forward(spec, data, targets, encoder_sequence=None, missing_mode="row", train=False,
            dropout_seed=None, row_offset=0, global_batch=None, keep_cache=False):
    """One batch through the chain.  Returns a dict of everything the kernel emits.

    data: list of (B, F_i) arrays indexed by POSITION in the sequence (multimodn.py:162-163);
    targets: (B, D) int or None (predict / get_states).
    global_batch: divisor for the means (data-parallel shards pass the global B).
    """
    S = spec["state_size"]
    E, D = len(spec["encoders"]), len(spec["decoders"])
    dtype = spec["init_state"].dtype
    B = data[0].shape[0]
    Bg = float(global_batch if global_batch is not None else B)
    seq = resolve_sequence(encoder_sequence, E)
    enc_layers = [expand_encoder(e, S) for e in spec["encoders"]]
    dec_layers = [expand_decoder(d) for d in spec["decoders"]]

    rnd = _rounding(spec)
    ce = np.zeros((E + 1, D))
    n_correct = np.zeros((E + 1, D))
    cm = np.zeros((4, E + 1, D))           # tp, tn, fp, fn
    n_present = np.zeros(E + 1)
    visited = np.zeros(E + 1, dtype=bool)
    sc = np.zeros(E)
    preds = np.zeros((E + 1, D, B), dtype=np.int64)
    outputs = [[None] * D for _ in range(E + 1)]
    cache = dict(steps=[], dec=[[None] * D for _ in range(E + 1)]) if keep_cache else None

    def eval_decoders(state, row, mask):
        for d in range(D):
            p, dcache = _mlp_forward(dec_layers[d], state, None, None, rnd, round_last=False)
            pred = _first_argmax(p)
            preds[row, d] = pred
            outputs[row][d] = p
            if keep_cache:
                cache["dec"][row][d] = (dcache, mask)
            if targets is None:
                continue
            y = targets[:, d].astype(np.int64)
            ce[row, d] = float((_ce_rows(p, y).astype(np.float64) * mask).sum() / Bg)
            n_correct[row, d] += float(((pred == y) * mask).sum())
            if spec["decoders"][d]["n_classes"] == 2:       # multimodn.py:153-157
                cm[0, row, d] += float(((pred == 1) & (y == 1) & mask).sum())
                cm[1, row, d] += float(((pred == 0) & (y == 0) & mask).sum())
                cm[2, row, d] += float(((pred == 1) & (y == 0) & mask).sum())
                cm[3, row, d] += float(((pred == 0) & (y == 1) & mask).sum())
            else:                                           # multimodn.py:60-63
                cm[:, row, d] = np.nan

    state = np.tile(spec["init_state"][None, :], (B, 1))     # state.py:30
    if rnd is not None:
        state = rnd(state)
    states = [state]
    ones = np.ones(B, dtype=bool)
    n_present[0] = B
    visited[0] = True
    eval_decoders(state, 0, ones)

    for pos, e in seq:
        x = np.asarray(data[pos]).astype(dtype)
        nan = np.isnan(x)
        if missing_mode == "batch":                          # multimodn.py:167-169
            present = np.zeros(B, dtype=bool) if nan.any() else ones
        elif missing_mode == "row":
            present = ~nan.any(axis=1)
        else:
            raise ValueError(missing_mode)
        step = dict(pos=pos, enc=e, present=present, s_prev=state)
        if present.any():
            xc = np.where(nan, 0, x)
            keep_scale = None
            enc = spec["encoders"][e]
            p_drop = float(enc.get("dropout", 0.0))
            if train and enc["kind"] == "mimic" and p_drop > 0:
                keep = dropout_keep(dropout_seed, e, np.arange(B) + row_offset, x.shape[1] + S, p_drop)
                keep_scale = (keep * dtype.type(1.0 / (1.0 - p_drop))).astype(dtype)
            s_hat, ecache = _mlp_forward(enc_layers[e], xc, state, keep_scale, rnd)
            new_state = np.where(present[:, None], s_hat, state)
            sc[e] = float((((new_state - state).astype(np.float64)) ** 2).sum() / (Bg * S))  # :174
            step.update(cache=ecache, keep_scale=keep_scale)
            state = new_state
            n_present[e + 1] += present.sum()                # :171
            visited[e + 1] = True
            eval_decoders(state, e + 1, present)
        else:
            # whole step skipped: history row e+1 untouched; predictions = decoders on the
            # unchanged state (deliberate deviation B2: reference predict has no NaN check)
            for d in range(D):
                p, _ = _mlp_forward(dec_layers[d], state, None, None, rnd, round_last=False)
                preds[e + 1, d] = _first_argmax(p)
                outputs[e + 1][d] = p
        step["s_new"] = state
        states.append(state)
        if keep_cache:
            cache["steps"].append(step)

    return dict(ce=ce, n_correct=n_correct, tp=cm[0], tn=cm[1], fp=cm[2], fn=cm[3],
                n_present=n_present, visited=visited, state_change=sc, predictions=preds,
                outputs=outputs, states=states, final_state=state, seq=seq, cache=cache,
                enc_layers=enc_layers, dec_layers=dec_layers)


def loss_from(fwd, spec, err_penalty, state_change_penalty_scaled):
    """multimodn.py:194-202 (state_change_penalty_scaled already includes the 0.01 of :86)."""
    E, D = len(spec["encoders"]), len(spec["decoders"])
    return (fwd["ce"].sum() / (D * (E + 1)) * err_penalty
            + fwd["state_change"].sum() / E * state_change_penalty_scaled)


# --------------------------------------------------------------------------------------
# backward (hand-derived)
# --------------------------------------------------------------------------------------
def _mlp_backward(layers, cache, dout, S, grads, keep_scale=None, rnd=None):
    """returns (d_first_input_without_state_or_None, d_state_or_None); accumulates grads."""
    d_state = None
    da = dout
    for j in range(len(layers) - 1, -1, -1):
        L = layers[j]
        inp, out = cache[j]
        dz = act_bwd(L["act"], out, da)
        W = L["W"]
        if rnd is not None:
            dz, W = rnd(dz), rnd(W)
        grads[j][0] += dz.T @ inp
        grads[j][1] += dz.sum(axis=0)
        dinp = dz @ W
        if j == 0 and keep_scale is not None:
            dinp = dinp * keep_scale
        if L["has_state"]:
            d_state = dinp[:, L["in_dim"]:]
            da = dinp[:, :L["in_dim"]]
        else:
            da = dinp
    return da, d_state


def train_step(spec, data, targets, err_penalty, state_change_penalty_scaled, encoder_sequence=None,
               missing_mode="row", dropout_seed=None, row_offset=0, global_batch=None, train=True):
    """forward + hand-written backward.  Returns (fwd dict, loss, grads dict, touched flags).

    grads: {'init_state': (S,), 'encoders': [[(dW, db), ...]], 'decoders': [[(dW, db), ...]]}
    touched[e] is False when encoder e received no gradient (reference: ``.grad is None``,
    multimodn.py:137,168-169).
    """
    S = spec["state_size"]
    E, D = len(spec["encoders"]), len(spec["decoders"])
    dtype = spec["init_state"].dtype
    B = data[0].shape[0]
    Bg = float(global_batch if global_batch is not None else B)
    fwd = forward(spec, data, targets, encoder_sequence, missing_mode, train=train,
                  dropout_seed=dropout_seed, row_offset=row_offset, global_batch=global_batch,
                  keep_cache=True)
    loss = loss_from(fwd, spec, err_penalty, state_change_penalty_scaled)
    enc_layers, dec_layers = fwd["enc_layers"], fwd["dec_layers"]
    g_enc = [[[np.zeros_like(L["W"]), np.zeros_like(L["b"])] for L in ls] for ls in enc_layers]
    g_dec = [[[np.zeros_like(L["W"]), np.zeros_like(L["b"])] for L in ls] for ls in dec_layers]
    touched = np.zeros(E, dtype=bool)
    rnd = _rounding(spec)
    c_err = dtype.type(err_penalty / (D * (E + 1) * Bg))
    c_sc = dtype.type(2.0 * state_change_penalty_scaled / (E * Bg * S))

    def decoders_backward(row):
        g = 0
        for d in range(D):
            entry = fwd["cache"]["dec"][row][d]
            if entry is None:
                continue
            dcache, mask = entry
            p = dcache[-1][1]
            y = targets[:, d].astype(np.int64)
            m = p.max(axis=1, keepdims=True)
            ex = np.exp(p - m)
            sm = ex / ex.sum(axis=1, keepdims=True)
            sm[np.arange(B), y] -= 1
            dp = (sm * (c_err * mask.astype(dtype))[:, None]).astype(dtype)
            da, _ = _mlp_backward(dec_layers[d], dcache, dp, S, g_dec[d], rnd=rnd)
            g = g + da
        return g

    G = np.zeros((B, S), dtype=dtype)
    for step in reversed(fwd["cache"]["steps"]):
        e = step["enc"]
        present = step["present"]
        if not present.any():
            continue                       # skipped step: state unchanged, nothing to add
        G = G + decoders_backward(e + 1)
        u = c_sc * (step["s_new"] - step["s_prev"])
        G = G + u
        dout = G * present[:, None]
        _, d_state = _mlp_backward(enc_layers[e], step["cache"], dout, S, g_enc[e], step.get("keep_scale"), rnd=rnd)
        touched[e] = True
        G = np.where(present[:, None], d_state, G) - u
    G = G + decoders_backward(0)
    g_init = G.sum(axis=0)
    grads = dict(init_state=g_init,
                 encoders=[[(w, b) for w, b in ls] for ls in g_enc],
                 decoders=[[(w, b) for w, b in ls] for ls in g_dec])
    return fwd, loss, grads, touched


# --------------------------------------------------------------------------------------
# epoch-level bookkeeping (history matrices)
# --------------------------------------------------------------------------------------
class EpochAccumulator:
    """multimodn.py:104-115 / 269-280 accumulators and :222-250 / :367-409 finalisation."""

    def __init__(self, E, D):
        self.E, self.D = E, D
        self.n_batches = 0
        self.n_samples = np.ones((E + 1, 1))            # starts at ONE (multimodn.py:105,270)
        self.err = np.zeros((E + 1, D))
        self.sc = np.zeros(E)
        self.n_correct = np.zeros((E + 1, D))
        self.cm = np.zeros((4, E + 1, D))

    def add(self, fwd):
        self.n_batches += 1
        self.n_samples[:, 0] += fwd["n_present"]
        self.err += fwd["ce"]
        self.sc += fwd["state_change"]
        self.n_correct += fwd["n_correct"]
        self.cm += np.stack([fwd["tp"], fwd["tn"], fwd["fp"], fwd["fn"]])

    def finalize(self):
        tp, tn, fp, fn = self.cm
        with np.errstate(divide="ignore", invalid="ignore"):
            sens = np.where(tp + fn == 0, 0, tp / (tp + fn))
            spec_ = np.where(tn + fp == 0, 0, tn / (tn + fp))
        return dict(loss=self.err / self.n_batches, state_change=self.sc / self.n_batches,
                    accuracy=self.n_correct / self.n_samples, sensitivity=sens, specificity=spec_,
                    balanced_accuracy=(sens + spec_) / 2)


def predict(spec, data, encoder_sequence=None, missing_mode="row"):
    """multimodn.py:422-458 -> float64 (E+1, D, N) class ids."""
    fwd = forward(spec, data, None, encoder_sequence, missing_mode)
    return fwd["predictions"].astype(np.float64)


def get_states(spec, data, encoder_sequence=None, missing_mode="row"):
    """multimodn.py:460-492 -> (N, S) final states."""
    return forward(spec, data, None, encoder_sequence, missing_mode)["final_state"]


# --------------------------------------------------------------------------------------
# Adam (torch.optim.Adam defaults; used by the history-parity tests and the CPU baseline)
# --------------------------------------------------------------------------------------
class Adam:
    """torch.optim.Adam(params, lr) with default betas/eps, no weight decay, no amsgrad.
    A parameter whose gradient is None this step is left untouched, moments and step
    count included (what torch does for ``p.grad is None``)."""

    def __init__(self, lr, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
        self.state = {}

    def update(self, key, p, g):
        st = self.state.setdefault(key, dict(t=0, m=np.zeros_like(p), v=np.zeros_like(p)))
        st["t"] += 1
        t = st["t"]
        dt = p.dtype.type
        st["m"] = st["m"] * dt(self.b1) + g * dt(1 - self.b1)
        st["v"] = st["v"] * dt(self.b2) + g * g * dt(1 - self.b2)
        bc1 = 1 - self.b1 ** t
        bc2 = 1 - self.b2 ** t
        step_size = self.lr / bc1
        denom = np.sqrt(st["v"]) / dt(np.sqrt(bc2)) + dt(self.eps)
        return (p - dt(step_size) * (st["m"] / denom)).astype(p.dtype)


def apply_adam(spec, grads, touched, opt):
    spec["init_state"] = opt.update("init", spec["init_state"], grads["init_state"].astype(spec["init_state"].dtype))
    for e, enc in enumerate(spec["encoders"]):
        if not touched[e]:
            continue
        enc["layers"] = [(opt.update(("e", e, j, 0), W, grads["encoders"][e][j][0]),
                          opt.update(("e", e, j, 1), b, grads["encoders"][e][j][1]))
                         for j, (W, b) in enumerate(enc["layers"])]
    for d, dec in enumerate(spec["decoders"]):
        dec["layers"] = [(opt.update(("d", d, j, 0), W, grads["decoders"][d][j][0]),
                          opt.update(("d", d, j, 1), b, grads["decoders"][d][j][1]))
                         for j, (W, b) in enumerate(dec["layers"])]
    return spec
