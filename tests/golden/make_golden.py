#!/usr/bin/env python
"""Generate the golden vectors in this directory by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, which does not exist on the GPU
box).  The .npz files it writes are committed; the tests only ever read those files.

    python tests/golden/make_golden.py

Every fixture stores: the model weights (``spec_*``), the inputs, and what the reference
produced on them (history matrices, gradients, predictions, states).  The reference imports
torchmetrics / torchsummary / matplotlib, which are absent here; ``oracle/ref_shims`` stubs
them (only binary ConfusionMatrix does arithmetic, everything else returns NaN).
"""
import itertools
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
torch._utils._accumulate = itertools.accumulate          # removed from modern torch (datasets/multimod_dataset.py:6)

import torch.nn.functional as F  # noqa: E402
from multimodn.multimodn import MultiModN  # noqa: E402  (the reference)
from multimodn.encoders import MLPEncoder, MIMIC_MLPEncoder, SLPEncoder, LinearEncoder, LogisticEncoder  # noqa: E402
from multimodn.decoders import LogisticDecoder, MLPDecoder, ClassDecoder  # noqa: E402
from multimodn.history import MultiModNHistory  # noqa: E402
from torch.nn import CrossEntropyLoss  # noqa: E402

from oracle.spec_io import spec_from_modules, spec_to_arrays, grads_to_arrays  # noqa: E402
from oracle.multimodn_oracle import dropout_keep  # noqa: E402

torch.set_num_threads(1)
CPU = torch.device("cpu")


class GradTap(torch.optim.Optimizer):
    """Optimizer that never moves the parameters; it sums the gradients it is handed and
    counts, per parameter, how often the gradient was None (skipped encoders)."""

    def __init__(self, params):
        super().__init__(list(params), {})
        self.acc = {}
        self.n_none = {}
        self.n_steps = 0

    def step(self):
        self.n_steps += 1
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    self.n_none[p] = self.n_none.get(p, 0) + 1
                else:
                    self.acc[p] = self.acc.get(p, 0) + p.grad.detach().clone()


def tapped_grads(model, tap, scale=1.0):
    def g(p):
        v = tap.acc.get(p)
        return (torch.zeros_like(p) if v is None else v * scale).numpy()

    grads = dict(init_state=g(model.init_state.state_value).reshape(-1), encoders=[], decoders=[])
    touched = []
    for enc in model.encoders:
        lin = [m for m in enc.layers if isinstance(m, torch.nn.Linear)]
        grads["encoders"].append([(g(l.weight), g(l.bias)) for l in lin])
        touched.append(lin[0].weight in tap.acc)
    for dec in model.decoders:
        lin = [dec.fc] if hasattr(dec, "fc") else list(dec.layers)
        grads["decoders"].append([(g(l.weight), g(l.bias)) for l in lin])
    return grads, np.array(touched)


def hist_arrays(history, tag, prefix):
    out = {}
    for name in ("loss", "accuracy", "sensitivity", "specificity", "balanced_accuracy"):
        out[f"{prefix}_{name}"] = np.stack([np.asarray(a, dtype=np.float64) for a in getattr(history, name)[tag]])
    if tag == "train":
        out[f"{prefix}_state_change"] = np.stack([np.asarray(a, dtype=np.float64) for a in history.state_change_loss])
    return out


def batches(data, y, bs, seq=None):
    out = []
    for i in range(0, y.shape[0], bs):
        item = [[torch.from_numpy(x[i:i + bs]) for x in data], torch.from_numpy(y[i:i + bs])]
        if seq is not None:
            item.append(torch.from_numpy(np.tile(np.asarray(seq)[None, :], (len(y[i:i + bs]), 1))))
        out.append(tuple(item))
    return out


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.1f} KiB")


def mimic_model(S, feats, hidden, D, dropout, err, scp, dec_hidden=None):
    encs = [MIMIC_MLPEncoder(S, f, hidden, activation=F.relu, dropout=dropout) for f in feats]
    decs = [MLPDecoder(S, dec_hidden if dec_hidden is not None else hidden, 2, output_activation=torch.sigmoid)
            for _ in range(D)]
    return MultiModN(S, encs, decs, err, scp, device=CPU)


# ---------------------------------------------------------------------------------------
def fixture_c1_titanic():
    """Config 1: pipelines/titanic/titanic_mlp_pipeline.py:26-85 on a synthetic Titanic-shaped
    table: 2 epochs of train_epoch (Adam lr 0.01) + test('val') per epoch; short last batch."""
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    n_train, n_val = 80, 40                               # 80 = 2 x 32 + 16: ragged last batch
    data = [rng.standard_normal((n_train, 6)).astype(np.float32)]
    y = (rng.random((n_train, 1)) < 0.4).astype(np.int64)
    vdata = [rng.standard_normal((n_val, 6)).astype(np.float32)]
    vy = (rng.random((n_val, 1)) < 0.4).astype(np.int64)
    model = MultiModN(1, [MLPEncoder(1, 6, (5, 5), F.relu)], [LogisticDecoder(1)], 0.7, 0.3, device=CPU)
    arrays = spec_to_arrays(spec_from_modules(model), "spec0")
    # single-batch gradients at the initial weights
    tap = GradTap(model.parameters())
    model.train_epoch(batches(data, y, 32)[:1], tap, CrossEntropyLoss())
    g, touched = tapped_grads(model, tap)
    arrays.update(grads_to_arrays(g, "grad0"))
    opt = torch.optim.Adam(list(model.parameters()), 0.01)
    hist = MultiModNHistory(["Survived"])
    for _ in range(2):
        model.train_epoch(batches(data, y, 32), opt, CrossEntropyLoss(), hist)
        model.test(batches(vdata, vy, 32), CrossEntropyLoss(), hist, tag="val")
    arrays.update(spec_to_arrays(spec_from_modules(model), "spec2"))
    arrays.update(hist_arrays(hist, "train", "train"))
    arrays.update(hist_arrays(hist, "val", "val"))
    arrays["predict"] = model.predict([torch.from_numpy(vdata[0])])
    arrays["states"] = torch.stack(model.get_states(batches(vdata, vy, 32))).numpy()
    save("c1_titanic", x0=data[0], y=y, vx0=vdata[0], vy=vy, err_penalty=0.7, state_change_penalty=0.3,
         lr=0.01, batch_size=32, **arrays)


def fixture_c2(name, S, feats, hidden, B, seed):
    """Config 2 (MIMIC-shaped): pipelines/mimic/mimic_multi_task_pipeline.py:118-120 model,
    dropout 0; one batch: gradients, train history, test history, predict, states."""
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    data = [rng.standard_normal((B, f)).astype(np.float32) for f in feats]
    y = (rng.random((B, 2)) < 0.3).astype(np.int64)
    model = mimic_model(S, feats, hidden, 2, 0.0, 1.0, 0.3)
    arrays = spec_to_arrays(spec_from_modules(model), "spec0")
    tap = GradTap(model.parameters())
    hist = MultiModNHistory(["a", "b"])
    model.train_epoch(batches(data, y, B), tap, CrossEntropyLoss(), hist)
    model.test(batches(data, y, B), CrossEntropyLoss(), hist, tag="val")
    g, _ = tapped_grads(model, tap)
    arrays.update(grads_to_arrays(g, "grad0"))
    arrays.update(hist_arrays(hist, "train", "train"))
    arrays.update(hist_arrays(hist, "val", "val"))
    arrays["predict"] = model.predict([torch.from_numpy(x) for x in data])
    arrays["states"] = torch.stack(model.get_states(batches(data, y, B))).numpy()
    # three Adam steps on the same batch (optimizer parity)
    opt = torch.optim.Adam(list(model.parameters()), 1e-3)
    for _ in range(3):
        model.train_epoch(batches(data, y, B), opt, CrossEntropyLoss())
    arrays.update(spec_to_arrays(spec_from_modules(model), "spec3"))
    save(name, y=y, err_penalty=1.0, state_change_penalty=0.3, lr=1e-3,
         **{f"x{i}": x for i, x in enumerate(data)}, **arrays)


def fixture_missing_row():
    """Row-level missingness == the reference run one row at a time (batch size 1, the
    reference's own recipe: pipelines/titanic/titanic_missingness_pipeline.py:35) and
    averaged.  Gradients are the mean of the per-row gradients."""
    torch.manual_seed(5)
    rng = np.random.default_rng(5)
    S, feats, B = 12, [5, 9, 14, 3], 24
    data = [rng.standard_normal((B, f)).astype(np.float32) for f in feats]
    y = (rng.random((B, 2)) < 0.5).astype(np.int64)
    for i, x in enumerate(data):
        miss = rng.random(B) < (0.5 if i != 2 else 0.25)
        x[miss, :] = np.nan
    data[1][:, :] = np.nan                                 # one modality missing for every row
    data[3][0, :] = 0.25                                   # reference test() needs the last encoder in batch 0
    data[0][3, :] = 0.5
    data[0][3, 2] = np.nan                                 # a single NaN feature also marks the row
    model = mimic_model(S, feats, (7, 6), 2, 0.0, 0.9, 0.4, dec_hidden=(5,))
    arrays = spec_to_arrays(spec_from_modules(model), "spec0")
    tap = GradTap(model.parameters())
    hist = MultiModNHistory(["a", "b"])
    model.train_epoch(batches(data, y, 1), tap, CrossEntropyLoss(), hist)
    try:
        # the history is appended (multimodn.py:390-409) before the end-of-test metric suite
        # trips over outputs/targets of different lengths (rows that skipped the last encoder)
        model.test(batches(data, y, 1), CrossEntropyLoss(), hist, tag="val")
    except Exception as exc:  # noqa: BLE001
        print("  reference test() raised after writing history:", type(exc).__name__)
    g, touched = tapped_grads(model, tap, scale=1.0 / B)
    arrays.update(grads_to_arrays(g, "grad0"))
    arrays.update(hist_arrays(hist, "train", "train"))
    arrays.update(hist_arrays(hist, "val", "val"))
    arrays["touched"] = touched
    arrays["states"] = torch.stack(model.get_states(batches(data, y, 1))).numpy()
    save("missing_row", y=y, err_penalty=0.9, state_change_penalty=0.4,
         **{f"x{i}": x for i, x in enumerate(data)}, **arrays)


def fixture_missing_batch():
    """Reference batch-level rule: one NaN anywhere skips that encoder for the whole batch
    (multimodn.py:167-169); skipped encoder => .grad is None."""
    torch.manual_seed(6)
    rng = np.random.default_rng(6)
    S, feats, B = 10, [4, 7, 6], 16
    data = [rng.standard_normal((2 * B, f)).astype(np.float32) for f in feats]
    y = (rng.random((2 * B, 2)) < 0.5).astype(np.int64)
    data[1][5, 3] = np.nan                                 # batch 0: encoder 1 skipped
    data[2][B + 2, :] = np.nan                             # batch 1: encoder 2 skipped
    model = mimic_model(S, feats, (8,), 2, 0.0, 1.0, 0.5)
    arrays = spec_to_arrays(spec_from_modules(model), "spec0")
    tap = GradTap(model.parameters())
    hist = MultiModNHistory(["a", "b"])
    model.train_epoch(batches(data, y, B)[:1], tap, CrossEntropyLoss(), hist)
    g, touched = tapped_grads(model, tap)
    arrays.update(grads_to_arrays(g, "grad0"))
    arrays["touched"] = touched
    hist2 = MultiModNHistory(["a", "b"])
    tap2 = GradTap(model.parameters())
    model.train_epoch(batches(data, y, B), tap2, CrossEntropyLoss(), hist2)
    try:   # as in fixture_missing_row: history is written before the metric suite raises
        model.test(batches(data, y, B), CrossEntropyLoss(), hist2, tag="val")
    except Exception as exc:  # noqa: BLE001
        print("  reference test() raised after writing history:", type(exc).__name__)
    arrays.update(hist_arrays(hist, "train", "train1"))
    arrays.update(hist_arrays(hist2, "train", "train2"))
    arrays.update(hist_arrays(hist2, "val", "val2"))
    arrays["states"] = torch.stack(model.get_states(batches(data, y, B))).numpy()
    save("missing_batch", y=y, err_penalty=1.0, state_change_penalty=0.5, batch_size=B,
         **{f"x{i}": x for i, x in enumerate(data)}, **arrays)


def fixture_sequence():
    """Permuted encoding_sequence (multimodn.py:509-531): history rows and predictions are
    indexed by ENCODER ID + 1, data by POSITION (multimodn.py:162-163,181,455)."""
    torch.manual_seed(7)
    rng = np.random.default_rng(7)
    S, B = 9, 20
    feats = [6, 6, 6, 6]                                   # same width so any order type-checks
    data = [rng.standard_normal((B, f)).astype(np.float32) for f in feats]
    y = (rng.random((B, 3)) < 0.5).astype(np.int64)
    seq = [2, 0, 3, 1]
    model = mimic_model(S, feats, (8, 8), 3, 0.0, 1.0, 1.0)
    arrays = spec_to_arrays(spec_from_modules(model), "spec0")
    tap = GradTap(model.parameters())
    hist = MultiModNHistory(["a", "b", "c"])
    model.train_epoch(batches(data, y, B, seq), tap, CrossEntropyLoss(), hist)
    model.test(batches(data, y, B, seq), CrossEntropyLoss(), hist, tag="val")
    g, _ = tapped_grads(model, tap)
    arrays.update(grads_to_arrays(g, "grad0"))
    arrays.update(hist_arrays(hist, "train", "train"))
    arrays.update(hist_arrays(hist, "val", "val"))
    arrays["predict"] = model.predict([torch.from_numpy(x) for x in data],
                                      torch.from_numpy(np.tile(np.array(seq)[None], (B, 1))))
    arrays["states"] = torch.stack(model.get_states(batches(data, y, B, seq))).numpy()
    save("sequence", y=y, seq=np.array(seq), err_penalty=1.0, state_change_penalty=1.0,
         **{f"x{i}": x for i, x in enumerate(data)}, **arrays)


def fixture_zoo():
    """Every dense encoder / decoder class of the reference in one model: MLPEncoder (sigmoid
    act), SLPEncoder, LinearEncoder, LogisticEncoder, MIMIC_MLPEncoder without hidden layers;
    ClassDecoder with 3 classes (non-binary => NaN confusion cells, multimodn.py:60-63),
    LogisticDecoder, MLPDecoder with 4 classes and tanh hidden activation."""
    torch.manual_seed(8)
    rng = np.random.default_rng(8)
    S, B = 7, 18
    feats = [5, 3, 4, 2, 6]
    data = [rng.standard_normal((B, f)).astype(np.float32) for f in feats]
    y = np.stack([rng.integers(0, 3, B), rng.integers(0, 2, B), rng.integers(0, 4, B)], axis=1).astype(np.int64)
    encs = [MLPEncoder(S, 5, (6, 4), torch.sigmoid), SLPEncoder(S, 3), LinearEncoder(S, 4),
            LogisticEncoder(S, 2), MIMIC_MLPEncoder(S, 6, (), dropout=0.0, activation=torch.tanh)]
    decs = [ClassDecoder(S, 3, torch.sigmoid), LogisticDecoder(S),
            MLPDecoder(S, (5,), 4, output_activation=torch.sigmoid, hidden_activation=torch.tanh)]
    model = MultiModN(S, encs, decs, 0.8, 2.0, device=CPU)
    arrays = spec_to_arrays(spec_from_modules(model), "spec0")
    tap = GradTap(model.parameters())
    hist = MultiModNHistory(["a", "b", "c"])
    model.train_epoch(batches(data, y, B), tap, CrossEntropyLoss(), hist)
    g, _ = tapped_grads(model, tap)
    arrays.update(grads_to_arrays(g, "grad0"))
    arrays.update(hist_arrays(hist, "train", "train"))
    arrays["predict"] = model.predict([torch.from_numpy(x) for x in data])
    arrays["states"] = torch.stack(model.get_states(batches(data, y, B))).numpy()
    save("zoo", y=y, err_penalty=0.8, state_change_penalty=2.0,
         **{f"x{i}": x for i, x in enumerate(data)}, **arrays)


def fixture_dropout():
    """Train-mode dropout of MIMIC_MLPEncoder (mlp_encoder.py:33-34,43-44) with the random
    stream replaced by the repo's counter-based mask (oracle.dropout_keep), so that everything
    downstream of the mask (scaling, forward, autograd) is the reference's own arithmetic."""
    torch.manual_seed(9)
    rng = np.random.default_rng(9)
    S, feats, B, p, seed = 8, [10, 21], 32, 0.25, 1234
    data = [rng.standard_normal((B, f)).astype(np.float32) for f in feats]
    y = (rng.random((B, 2)) < 0.5).astype(np.int64)
    model = mimic_model(S, feats, (8, 8), 2, p, 1.0, 0.3)
    for e, enc in enumerate(model.encoders):
        keep = torch.from_numpy(dropout_keep(seed, e, np.arange(B), feats[e] + S, p))
        scale = np.float32(1.0 / (1.0 - p))

        def fwd(x, keep=keep, scale=scale, enc=enc):
            return x * (keep * scale) if enc.training else x
        enc.layers[0].forward = fwd
    arrays = spec_to_arrays(spec_from_modules(model), "spec0")
    tap = GradTap(model.parameters())
    hist = MultiModNHistory(["a", "b"])
    model.train_epoch(batches(data, y, B), tap, CrossEntropyLoss(), hist)
    model.test(batches(data, y, B), CrossEntropyLoss(), hist, tag="val")
    g, _ = tapped_grads(model, tap)
    arrays.update(grads_to_arrays(g, "grad0"))
    arrays.update(hist_arrays(hist, "train", "train"))
    arrays.update(hist_arrays(hist, "val", "val"))
    save("dropout", y=y, err_penalty=1.0, state_change_penalty=0.3, dropout_seed=seed,
         **{f"x{i}": x for i, x in enumerate(data)}, **arrays)


def fixture_titanic_pipeline():
    """pipelines/titanic/titanic_mlp_pipeline.py:24-85 end to end with the reference's own
    TitanicDataset / PartitionDataset.random_split / DataLoader / MultiModN / MultiModNHistory.get_results,
    on the synthetic Titanic-shaped table of multimodn_b200/datasets/titanic.py (the real CSV is fetched by
    datasets/titanic/get_data.sh, impossible offline).  The reference hard-codes the CSV location inside its
    read-only tree (titanic_dataset.py:22), so pandas.read_csv is redirected for that one call."""
    import tempfile
    import pandas as pd
    from torch.utils.data import DataLoader
    import datasets.titanic.titanic_dataset as ref_td
    from multimodn_b200.datasets.titanic import write_synthetic_titanic_csv

    n_rows, seed, epochs, batch_size = 400, 0, 3, 32
    csv = os.path.join(tempfile.mkdtemp(), "titanic.csv")
    write_synthetic_titanic_csv(csv, n_rows, seed)
    real_read = pd.read_csv
    ref_td.pd.read_csv = lambda path, *a, **k: real_read(csv, *a, **k)
    try:
        features = ['Fare', 'Pclass', 'Age', 'Sex_male', 'Relatives', 'Embarked']
        targets = ['Survived']
        torch.manual_seed(seed)
        full = ref_td.TitanicDataset(features, targets, dropna=True, std=True)
        dataset = full.partition_dataset()
    finally:
        ref_td.pd.read_csv = real_read
    train_data, val_data, test_data = dataset.random_split((0.8, 0.2, 0), seed, 0)
    train_loader = DataLoader(train_data, batch_size)
    val_loader = DataLoader(val_data, batch_size)
    encoders = [MLPEncoder(1, len(features), (5, 5), F.relu)]
    decoders = [LogisticDecoder(1) for _ in targets]
    model = MultiModN(1, encoders, decoders, 0.7, 0.3, device=CPU)
    arrays = spec_to_arrays(spec_from_modules(model), "spec0")
    optimizer = torch.optim.Adam(list(model.parameters()), 0.01)
    history = MultiModNHistory(targets)
    for _ in range(epochs):
        model.train_epoch(train_loader, optimizer, CrossEntropyLoss(), history)
        model.test(val_loader, CrossEntropyLoss(), history, tag='val')
    arrays.update(spec_to_arrays(spec_from_modules(model), "spec_final"))
    arrays.update(hist_arrays(history, "train", "train"))
    arrays.update(hist_arrays(history, "val", "val"))
    results = history.get_results()
    save("titanic_pipeline", n_rows=n_rows, seed=seed, epochs=epochs, batch_size=batch_size,
         X=np.asarray(full.X, dtype=np.float64), y=np.asarray(full.y),
         train_idx=np.asarray(train_data.indices), val_idx=np.asarray(val_data.indices),
         test_idx=np.asarray(test_data.indices), results=results.to_numpy(dtype=np.float64),
         results_columns=np.array(list(results.columns)), results_index=np.array(list(results.index)), **arrays)


if __name__ == "__main__":
    fixture_c1_titanic()
    fixture_c2("c2_mimic_small", 16, [6, 19, 40], (8, 8), 16, 1)
    fixture_c2("c2_mimic_full", 64, [6, 99, 1024], (32, 32), 16, 1)
    fixture_missing_row()
    fixture_missing_batch()
    fixture_sequence()
    fixture_zoo()
    fixture_dropout()
    fixture_titanic_pipeline()
