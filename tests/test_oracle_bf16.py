"""CPU: the oracle's bf16 restatement (the yardstick of the wide regime's GPU tests) against the fp32 oracle:
per-step losses and history metrics within the north star's bf16 tolerance (1e-2), same skipped encoders, and
rounding that is idempotent and exact on bf16-representable values."""
import numpy as np
import torch

from oracle import multimodn_oracle as O
from oracle.spec_io import random_spec, synthetic_batch
from helpers import flat_grads, assert_close


def test_round_bf16_matches_torch():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(4096).astype(np.float32) * 10.0 ** rng.integers(-20, 20, 4096),
                        np.array([0.0, -0.0, 1.0, 1.00390625, 1.001953125, np.inf, -np.inf], dtype=np.float32)])
    want = torch.from_numpy(x).to(torch.bfloat16).float().numpy()
    got = O.round_bf16(x)
    assert (got == want).all()
    assert (O.round_bf16(got) == got).all()


def test_bf16_restatement_tracks_fp32_oracle():
    rng = np.random.default_rng(3)
    feats = [37, 70]
    spec = random_spec(rng, 60, feats, enc_kind="mimic", enc_hidden=(90, 77), dropout=0.2, n_decoders=2, dec_hidden=(48,), n_classes=2)
    data, y = synthetic_batch(rng, feats, 2, 300, mnar=True)
    s32 = O.cast_spec(spec, np.float32)
    sb = dict(s32, precision="bf16")
    f1, l1, g1, t1 = O.train_step(s32, data, y, 0.8, 0.006, dropout_seed=5)
    f2, l2, g2, t2 = O.train_step(sb, data, y, 0.8, 0.006, dropout_seed=5)
    assert (t1 == t2).all()
    assert abs(l1 - l2) <= 1e-2 * abs(l1)
    assert_close(f2["ce"], f1["ce"], rtol=1e-2, what="per-step losses")
    assert_close(f2["state_change"], f1["state_change"], rtol=1e-2, what="state change")
    a, b = flat_grads(g1).astype(np.float64), flat_grads(g2).astype(np.float64)
    assert a @ b / (np.linalg.norm(a) * np.linalg.norm(b)) > 0.999
    assert (f1["predictions"] != f2["predictions"]).mean() < 0.03
    assert O.cast_spec(sb, np.float32)["precision"] == "bf16"
