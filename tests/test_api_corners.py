"""CPU (emulated kernels): corners of the reference API that the golden fixtures do not reach —
``shuffle_mode`` (multimodn.py:527-529), ``last_epoch`` (:251-252), ``log_interval`` (:214-220) and the
CrossEntropyLoss contract on out-of-range targets."""
import random

import numpy as np
import pytest
import torch
from torch.nn import CrossEntropyLoss

from oracle import multimodn_oracle as O
from oracle.spec_io import random_spec, synthetic_batch
from multimodn_b200 import MultiModNHistory
from helpers import flat_grads, assert_close
from model_utils import model_from_spec, GradTap, tapped_flat


def _setup(shuffle_mode=False, B=70, seed=3):
    rng = np.random.default_rng(seed)
    feats = [5, 9, 14]
    spec = random_spec(rng, 12, feats, enc_hidden=(8,), n_decoders=2, dec_hidden=(8,))
    data, y = synthetic_batch(rng, feats, 2, B, mnar=True)
    model = model_from_spec(spec, 0.8, 0.6, "cpu", "row", shuffle_mode=shuffle_mode)
    loader = [([torch.from_numpy(x) for x in data], torch.from_numpy(y))]
    return spec, data, y, model, loader


def test_shuffle_mode_trains_on_the_shuffled_sequence(emu):
    """train=True shuffles the (position, encoder) pairs with Python's `random` (multimodn.py:527-529); the step must
    equal the oracle on exactly that order, and eval-mode calls must keep the natural order."""
    spec, data, y, model, loader = _setup(shuffle_mode=True)
    tap = GradTap(model.parameters())
    hist = MultiModNHistory(["a", "b"])
    random.seed(1234)
    pairs = list(enumerate(range(3)))
    random.shuffle(pairs)                       # what get_encoder_iterable will draw
    assert pairs != list(enumerate(range(3))), "pick a seed that actually permutes"
    random.seed(1234)
    model.train_epoch(loader, tap, CrossEntropyLoss(), hist)
    got, _ = tapped_flat(model, tap)
    # the oracle takes data indexed by POSITION in the sequence and a sequence of encoder ids
    odata = [data[p] for p, _ in pairs]
    oseq = [e for _, e in pairs]
    fwd, _, grads, _ = O.train_step(O.cast_spec(spec, np.float32), odata, y, 0.8, 0.006, encoder_sequence=oseq)
    assert_close(got, flat_grads(grads), rtol=1e-5, what="grads under shuffle_mode")
    assert_close(hist.loss["train"][0], fwd["ce"], rtol=1e-5, what="loss under shuffle_mode")
    # eval: no shuffle
    assert model.get_encoder_iterable(None, True, train=False) == list(enumerate(range(3)))
    pred = model.predict([torch.from_numpy(x) for x in data])
    assert (pred == O.forward(O.cast_spec(spec, np.float32), data, y)["predictions"]).mean() > 0.998


def test_last_epoch_returns_test_on_the_train_loader(emu):
    spec, data, y, model, loader = _setup()
    tap = GradTap(model.parameters())
    hist = MultiModNHistory(["a", "b"])
    res = model.train_epoch(loader, tap, CrossEntropyLoss(), hist, last_epoch=True)
    want = model.test(loader, CrossEntropyLoss())
    assert isinstance(res, list) and len(res) == 2 and len(res) == len(want)
    assert len(hist.loss["train"]) == 1 and "test" not in hist.loss or len(hist.loss.get("test", [])) == 0   # history=None inside
    assert model.train_epoch(loader, tap, CrossEntropyLoss(), hist) is None


def test_log_interval_reports_each_batch_loss(emu):
    spec, data, y, model, _ = _setup(B=96)
    bs = 32
    loader = [([torch.from_numpy(x[i:i + bs]) for x in data], torch.from_numpy(y[i:i + bs])) for i in range(0, 96, bs)]
    tap = GradTap(model.parameters())
    lines = []
    hist = MultiModNHistory(["a", "b"])
    model.train_epoch(loader, tap, CrossEntropyLoss(), hist, log_interval=2, logger=lines.append)
    assert len(lines) == 1 and lines[0].startswith("Batch 2/3")          # batch_idx % 2 == 1 only (multimodn.py:214)
    s32 = O.cast_spec(spec, np.float32)
    fwd = O.forward(s32, [x[32:64] for x in data], y[32:64], None, "row", train=True)
    loss = O.loss_from(fwd, s32, 0.8, 0.006)
    got = float(lines[0].split("Loss: ")[1].split("\n")[0])
    assert abs(got - loss) <= 1e-4 * max(1.0, abs(loss)) + 5e-5         # printed with 4 decimals
    # the epoch history is unaffected by logging
    acc = O.EpochAccumulator(3, 2)
    for i in range(0, 96, bs):
        acc.add(O.forward(s32, [x[i:i + bs] for x in data], y[i:i + bs], None, "row", train=True))
    assert_close(hist.loss["train"][0], acc.finalize()["loss"], rtol=1e-5, what="epoch loss with log_interval")


@pytest.mark.parametrize("bad", [2, -1, -100])
def test_out_of_range_target_raises_like_cross_entropy(emu, bad):
    """nn.CrossEntropyLoss raises on a target outside [0, C) (the reference would, multimodn.py:146); the fused step must
    not silently clamp it"""
    spec, data, y, model, _ = _setup()
    y = y.copy()
    y[5, 1] = bad
    loader = [([torch.from_numpy(x) for x in data], torch.from_numpy(y))]
    tap = GradTap(model.parameters())
    with pytest.raises(IndexError, match="Target out of bounds"):
        model.train_epoch(loader, tap, CrossEntropyLoss(), MultiModNHistory(["a", "b"]))
    # the flag is cleared: a clean epoch afterwards passes
    y[5, 1] = 1
    loader = [([torch.from_numpy(x) for x in data], torch.from_numpy(y))]
    model.test(loader, CrossEntropyLoss(), MultiModNHistory(["a", "b"]))
