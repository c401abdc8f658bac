"""Parity cases shared by the CPU (emulated kernels) and GPU (libmmn.so) suites: product path vs
the oracle on seeded synthetic inputs, including multi-tile batches, ragged last tiles, every
tile height (RM = 4, 2, 1), MNAR missingness and dropout."""
import numpy as np
import torch
from torch.nn import CrossEntropyLoss

from oracle import multimodn_oracle as O
from oracle.spec_io import random_spec, synthetic_batch
from multimodn_b200 import MultiModNHistory
from helpers import flat_grads, assert_close
from model_utils import model_from_spec, GradTap, tapped_flat

HIST = ("loss", "accuracy", "sensitivity", "specificity", "balanced_accuracy")

# name: (S, features, enc_kind, enc_hidden, D, dec_hidden, n_classes, B, mnar, dropout)
CASES = {
    "multi_tile_ragged": (16, [5, 12, 33], "mimic", (8, 8), 2, (8,), 2, 3 * 128 * 3 + 37, False, 0.0),
    "mnar_rows": (24, [6, 9, 17, 4], "mimic", (16,), 3, (8, 8), 2, 517, True, 0.0),
    "mlp_kind_wide_hidden": (8, [7, 40], "mlp", (40, 12), 2, (), 2, 300, True, 0.0),
    "rm2_state300": (300, [10, 6], "mimic", (8,), 1, (8,), 2, 150, False, 0.0),
    "rm1_state600": (600, [9], "mimic", (4,), 1, (), 3, 70, True, 0.0),
    "dropout_mnar": (16, [20, 9], "mimic", (8, 8), 2, (8,), 2, 400, True, 0.3),
    "hidden_wider_than_32": (12, [50], "mimic", (70, 33), 2, (40,), 5, 200, False, 0.0),
}


def run_parity_case(name, device, rtol=1e-5, missing_mode="row"):
    S, feats, kind, eh, D, dh, C, B, mnar, p = CASES[name]
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
    rng = np.random.default_rng(sum(map(ord, name)))
    spec = random_spec(rng, S, feats, enc_kind=kind, enc_hidden=eh, dropout=p, n_decoders=D, dec_hidden=dh, n_classes=C)
    data, y = synthetic_batch(rng, feats, D, B, mnar=mnar, n_classes=C)
    err, scp = 0.8, 0.6
    model = model_from_spec(spec, err, scp, device, missing_mode)
    tap = GradTap(model.parameters())
    hist = MultiModNHistory([str(i) for i in range(D)])
    loader = [([torch.from_numpy(x).to(device) for x in data], torch.from_numpy(y).to(device))]
    rt = model.runtime()
    rt.dropout_base_seed, rt.step_counter = 77, 0
    seed = (77 * 0x9E3779B1 + 1 * 0x85EBCA77) & 0xFFFFFFFF
    model.train_epoch(loader, tap, CrossEntropyLoss(), hist)
    got, touched = tapped_flat(model, tap)
    fwd, loss, grads, otouched = O.train_step(O.cast_spec(spec, np.float32), data, y, err, 0.01 * scp,
                                              missing_mode=missing_mode, dropout_seed=seed)
    assert (touched == otouched).all()
    assert_close(got, flat_grads(grads), rtol=rtol, what=f"{name}: grads")
    acc = O.EpochAccumulator(len(feats), D)
    acc.add(fwd)
    fin = acc.finalize()
    for n in HIST:
        assert_close(getattr(hist, n)["train"][0], fin[n], rtol=rtol, what=f"{name}: train {n}")
    assert_close(hist.state_change_loss[0], fin["state_change"], rtol=rtol, what=f"{name}: state_change")
    # eval-mode forward: predictions bit-exact, states and missing-row pass-through
    ofwd = O.forward(O.cast_spec(spec, np.float32), data, y, None, missing_mode)
    pred = model.predict([torch.from_numpy(x) for x in data])
    mismatch = (pred != ofwd["predictions"]).mean()
    assert mismatch <= 2e-3, f"{name}: {mismatch:.4f} of predictions differ"   # arg-max ties at fp32 round-off only
    states = torch.stack(model.get_states(loader)).cpu().numpy()
    assert_close(states, ofwd["final_state"], rtol=rtol, what=f"{name}: states")
    assert not np.isnan(states).any()
    model.test(loader, CrossEntropyLoss(), hist, tag="val")
    acc = O.EpochAccumulator(len(feats), D)
    acc.add(ofwd)
    fin = acc.finalize()
    assert_close(hist.loss["val"][0], fin["loss"], rtol=rtol, what=f"{name}: val loss")
    return model
