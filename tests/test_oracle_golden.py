"""CPU: the oracle (oracle/multimodn_oracle.py) against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import copy

import numpy as np
import pytest

from oracle import multimodn_oracle as O
from helpers import (load_golden, golden_data, golden_spec, golden_grads, flat_grads, flat_params,
                     assert_close)

HIST = ("loss", "accuracy", "sensitivity", "specificity", "balanced_accuracy")


def epoch(spec, data, y, bs, seq=None, missing_mode="row", train=False, dropout_seed=None):
    E, D = len(spec["encoders"]), len(spec["decoders"])
    acc = O.EpochAccumulator(E, D)
    for i in range(0, len(y), bs):
        fwd = O.forward(spec, [x[i:i + bs] for x in data], y[i:i + bs], seq, missing_mode, train=train,
                        dropout_seed=dropout_seed)
        acc.add(fwd)
    return acc.finalize()


def check_history(fx, prefix, hist, idx=0, with_sc=False, rtol=2e-6):
    for name in HIST:
        assert_close(hist[name], fx[f"{prefix}_{name}"][idx], rtol=rtol, what=f"{prefix}_{name}")
    if with_sc:
        assert_close(hist["state_change"], fx[f"{prefix}_state_change"][idx], rtol=rtol, what="state_change")


@pytest.mark.parametrize("name", ["c2_mimic_small", "c2_mimic_full"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_c2_step(name, dtype):
    fx = load_golden(name)
    spec = O.cast_spec(golden_spec(fx), dtype)
    data, y = golden_data(fx), fx["y"]
    fwd, loss, grads, touched = O.train_step(spec, data, y, float(fx["err_penalty"]),
                                             0.01 * float(fx["state_change_penalty"]), missing_mode="batch")
    assert touched.all()
    tol = 1e-5 if dtype == np.float32 else 2e-6        # the fixture itself is fp32 arithmetic
    assert_close(flat_grads(grads), flat_grads(golden_grads(fx, spec)), rtol=tol, what="grads")
    acc = O.EpochAccumulator(3, 2)
    acc.add(fwd)
    check_history(fx, "train", acc.finalize(), with_sc=True, rtol=tol)
    check_history(fx, "val", epoch(spec, data, y, len(y)), rtol=tol)
    assert (O.predict(spec, data) == fx["predict"]).all()
    assert_close(O.get_states(spec, data), fx["states"], rtol=tol, what="states")


def test_c2_adam_three_steps():
    fx = load_golden("c2_mimic_small")
    spec = O.cast_spec(golden_spec(fx), np.float32)
    data, y = golden_data(fx), fx["y"]
    opt = O.Adam(float(fx["lr"]))
    for _ in range(3):
        _, _, grads, touched = O.train_step(spec, data, y, 1.0, 0.01 * 0.3, missing_mode="batch")
        spec = O.apply_adam(spec, grads, touched, opt)
    assert_close(flat_params(spec), flat_params(golden_spec(fx, "spec3")), rtol=1e-5, what="params after Adam")


def test_c1_titanic_two_epochs():
    """full pipeline bookkeeping: ragged last batch, unweighted batch mean, np.ones counters,
    Adam between batches (multimodn.py:104-250)."""
    fx = load_golden("c1_titanic")
    spec = O.cast_spec(golden_spec(fx), np.float32)
    x, y, vx, vy = [fx["x0"]], fx["y"], [fx["vx0"]], fx["vy"]
    bs = int(fx["batch_size"])
    _, _, grads, _ = O.train_step(spec, [x[0][:bs]], y[:bs], 0.7, 0.01 * 0.3, missing_mode="batch")
    assert_close(flat_grads(grads), flat_grads(golden_grads(fx, spec)), rtol=1e-5, what="grads")
    opt = O.Adam(float(fx["lr"]))
    for ep in range(2):
        acc = O.EpochAccumulator(1, 1)
        for i in range(0, len(y), bs):
            fwd, _, grads, touched = O.train_step(spec, [x[0][i:i + bs]], y[i:i + bs], 0.7, 0.01 * 0.3,
                                                  missing_mode="batch")
            acc.add(fwd)
            spec = O.apply_adam(spec, grads, touched, opt)
        check_history(fx, "train", acc.finalize(), idx=ep, with_sc=True, rtol=1e-5)
        check_history(fx, "val", epoch(spec, vx, vy, bs), idx=ep, rtol=1e-5)
    assert_close(flat_params(spec), flat_params(golden_spec(fx, "spec2")), rtol=1e-5, what="params")
    assert (O.predict(spec, vx) == fx["predict"]).all()
    assert_close(O.get_states(spec, vx), fx["states"], rtol=1e-5, what="states")


def test_missing_row_equals_reference_at_batch_size_one():
    fx = load_golden("missing_row")
    spec = O.cast_spec(golden_spec(fx), np.float32)
    data, y = golden_data(fx), fx["y"]
    fwd, loss, grads, touched = O.train_step(spec, data, y, float(fx["err_penalty"]),
                                             0.01 * float(fx["state_change_penalty"]), missing_mode="row")
    assert (touched == fx["touched"]).all() and not touched[1]
    assert_close(flat_grads(grads), flat_grads(golden_grads(fx, spec)), rtol=1e-5, what="grads")
    acc = O.EpochAccumulator(4, 2)
    acc.add(fwd)
    check_history(fx, "train", acc.finalize(), with_sc=True, rtol=1e-5)
    check_history(fx, "val", epoch(spec, data, y, len(y)), rtol=1e-5)
    # missing rows: state passes through bit-exactly
    states = O.get_states(spec, data)
    assert_close(states, fx["states"], rtol=1e-6, what="states")
    s3 = fwd["states"]
    for k, (pos, e) in enumerate(fwd["seq"]):
        absent = np.isnan(data[pos]).any(axis=1)
        assert (s3[k + 1][absent] == s3[k][absent]).all()
        assert not np.isnan(s3[k + 1]).any()


def test_missing_batch_reference_rule():
    fx = load_golden("missing_batch")
    spec = O.cast_spec(golden_spec(fx), np.float32)
    data, y = golden_data(fx), fx["y"]
    B = int(fx["batch_size"])
    fwd, loss, grads, touched = O.train_step(spec, [x[:B] for x in data], y[:B], 1.0, 0.01 * 0.5,
                                             missing_mode="batch")
    assert (touched == fx["touched"]).all() and list(touched) == [True, False, True]
    assert_close(flat_grads(grads), flat_grads(golden_grads(fx, spec)), rtol=1e-5, what="grads")
    acc = O.EpochAccumulator(3, 2)
    acc.add(fwd)
    check_history(fx, "train1", acc.finalize(), with_sc=True, rtol=1e-5)
    acc = O.EpochAccumulator(3, 2)
    for i in range(0, 2 * B, B):
        acc.add(O.train_step(spec, [x[i:i + B] for x in data], y[i:i + B], 1.0, 0.005, missing_mode="batch")[0])
    check_history(fx, "train2", acc.finalize(), with_sc=True, rtol=1e-5)
    check_history(fx, "val2", epoch(spec, data, y, B, missing_mode="batch"), rtol=1e-5)
    states = np.concatenate([O.get_states(spec, [x[i:i + B] for x in data], missing_mode="batch")
                             for i in range(0, 2 * B, B)])
    assert_close(states, fx["states"], rtol=1e-5, what="states")


def test_permuted_sequence():
    fx = load_golden("sequence")
    spec = O.cast_spec(golden_spec(fx), np.float32)
    data, y, seq = golden_data(fx), fx["y"], fx["seq"]
    fwd, loss, grads, touched = O.train_step(spec, data, y, 1.0, 0.01, encoder_sequence=seq, missing_mode="batch")
    assert_close(flat_grads(grads), flat_grads(golden_grads(fx, spec)), rtol=1e-5, what="grads")
    acc = O.EpochAccumulator(4, 3)
    acc.add(fwd)
    check_history(fx, "train", acc.finalize(), with_sc=True, rtol=1e-5)
    check_history(fx, "val", epoch(spec, data, y, len(y), seq=np.tile(seq[None], (len(y), 1))), rtol=1e-5)
    assert (O.predict(spec, data, seq) == fx["predict"]).all()
    assert_close(O.get_states(spec, data, seq), fx["states"], rtol=1e-5, what="states")
    with pytest.raises(ValueError):
        bad = np.tile(seq[None], (len(y), 1))
        bad[3] = bad[3][::-1]
        O.resolve_sequence(bad, 4)


def test_module_zoo():
    fx = load_golden("zoo")
    spec = O.cast_spec(golden_spec(fx), np.float32)
    data, y = golden_data(fx), fx["y"]
    fwd, loss, grads, touched = O.train_step(spec, data, y, float(fx["err_penalty"]),
                                             0.01 * float(fx["state_change_penalty"]), missing_mode="batch")
    assert_close(flat_grads(grads), flat_grads(golden_grads(fx, spec)), rtol=1e-5, what="grads")
    acc = O.EpochAccumulator(5, 3)
    acc.add(fwd)
    h = acc.finalize()
    check_history(fx, "train", h, with_sc=True, rtol=1e-5)
    assert np.isnan(h["sensitivity"][:, 0]).all() and not np.isnan(h["sensitivity"][:, 1]).any()
    assert (O.predict(spec, data) == fx["predict"]).all()
    assert_close(O.get_states(spec, data), fx["states"], rtol=1e-5, what="states")


def test_dropout_fixed_mask():
    fx = load_golden("dropout")
    spec = O.cast_spec(golden_spec(fx), np.float32)
    data, y = golden_data(fx), fx["y"]
    seed = int(fx["dropout_seed"])
    fwd, loss, grads, touched = O.train_step(spec, data, y, 1.0, 0.003, missing_mode="batch", dropout_seed=seed)
    assert_close(flat_grads(grads), flat_grads(golden_grads(fx, spec)), rtol=1e-5, what="grads")
    acc = O.EpochAccumulator(2, 2)
    acc.add(fwd)
    check_history(fx, "train", acc.finalize(), with_sc=True, rtol=1e-5)
    check_history(fx, "val", epoch(spec, data, y, len(y)), rtol=1e-5)       # eval: no dropout
    keep = O.dropout_keep(seed, 0, np.arange(4096), 64, 0.25)
    assert abs(keep.mean() - 0.75) < 0.01


def test_torch_port_matches_oracle():
    """the CPU-baseline port (oracle/torch_port.py, timed by bench.py) computes the same step"""
    import torch
    from oracle.torch_port import TorchPort
    from oracle.spec_io import random_spec, synthetic_batch
    rng = np.random.default_rng(11)
    feats = [6, 19, 40]
    spec = random_spec(rng, 16, feats, enc_hidden=(8, 8), n_decoders=2, dec_hidden=(8, 8))
    data, y = synthetic_batch(rng, feats, 2, 96, mnar=True)
    fwd, loss, grads, touched = O.train_step(O.cast_spec(spec, np.float32), data, y, 1.0, 0.003, missing_mode="row")
    port = TorchPort(spec, 1.0, 0.003)
    ploss, ce, sc, pg = port.grads([torch.from_numpy(x) for x in data], torch.from_numpy(y))
    assert abs(ploss - loss) <= 1e-5 * abs(loss)
    assert_close(ce, fwd["ce"], rtol=1e-5, what="ce")
    assert_close(sc, fwd["state_change"], rtol=1e-5, what="sc")
    assert_close(np.concatenate([g.ravel() for g in pg]), flat_grads(grads), rtol=1e-5, what="grads")
