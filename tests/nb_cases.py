"""Parity cases of the bf16 per-tile kernel (mmn_nb.cuh: precision="bf16" on narrow models), shared by the CPU suite
(kernels on the host emulator, lane-exact models of mma.sync / ldmatrix / movmatrix) and the GPU suite.

Two yardsticks, as for the wide regime (tests/test_gpu_wide.py):
  * the oracle's bf16 restatement — the same algorithm rounded to bfloat16 exactly where the kernel rounds (weights, layer
    inputs, activations, states, layer gradients; fp32 sums, biases, state gradient, parameter gradients): gradients, losses
    and states agree to summation-order accuracy;
  * the fp32 oracle (= the reference): per-step losses, history metrics and predictions within the north star's bf16
    tolerance of 1e-2 relative."""
import numpy as np
import torch
from torch.nn import CrossEntropyLoss

from oracle import multimodn_oracle as O
from oracle.spec_io import random_spec, synthetic_batch
from multimodn_b200 import MultiModNHistory
from helpers import flat_grads, assert_close
from model_utils import model_from_spec, GradTap, tapped_flat

HIST = ("loss", "accuracy", "sensitivity", "specificity", "balanced_accuracy")
ENGINE_NB = 4

# name: (S, features, enc_kind, enc_hidden, D, dec_hidden, n_classes, B, mnar, dropout)
CASES = {
    "c2_shape": (64, [6, 99, 256], "mimic", (32, 32), 2, (32, 32), 2, 300, False, 0.0),
    "ragged_dims_mnar": (24, [6, 9, 17, 4], "mimic", (20,), 3, (8, 12), 2, 517, True, 0.0),
    "mlp_kind": (8, [7, 40], "mlp", (40, 12), 2, (), 2, 300, True, 0.0),
    "slp_and_class_decoder": (12, [5, 33], "mlp", (), 2, (), 3, 200, True, 0.0),
    "dropout_mnar": (16, [20, 9], "mimic", (8, 8), 2, (8,), 2, 400, True, 0.3),
    "wide_hidden_multiclass": (48, [50], "mimic", (64, 33), 2, (40,), 5, 200, False, 0.0),
    "titanic_shape": (1, [6], "mlp", (5, 5), 1, (), 2, 150, False, 0.0),
}


def run_case(name, device, lib, missing_mode="row", B=None):
    S, feats, kind, eh, D, dh, C, B0, mnar, p = CASES[name]
    B = B or B0
    rng = np.random.default_rng(sum(map(ord, name)))
    spec = random_spec(rng, S, feats, enc_kind=kind, enc_hidden=eh, dropout=p, n_decoders=D, dec_hidden=dh, n_classes=C)
    data, y = synthetic_batch(rng, feats, D, B, mnar=mnar, n_classes=C)
    err, scp = 0.8, 0.6
    model = model_from_spec(spec, err, scp, device, missing_mode, precision="bf16")
    rt = model.runtime()
    assert lib.dll.mmn_plan_engine(rt.plan) == ENGINE_NB, "the model should qualify for the bf16 tile kernel"
    tap = GradTap(model.parameters())
    hist = MultiModNHistory([str(i) for i in range(D)])
    loader = [([torch.from_numpy(x).to(device) for x in data], torch.from_numpy(y).to(device))]
    rt.dropout_base_seed, rt.step_counter = 77, 0
    seed = (77 * 0x9E3779B1 + 1 * 0x85EBCA77) & 0xFFFFFFFF
    model.train_epoch(loader, tap, CrossEntropyLoss(), hist)
    got, touched = tapped_flat(model, tap)
    assert np.isfinite(got).all()
    pred = model.predict([torch.from_numpy(x) for x in data])
    states = torch.stack(model.get_states(loader)).cpu().numpy()
    model.test(loader, CrossEntropyLoss(), hist, tag="val")

    s32 = O.cast_spec(spec, np.float32)
    for yard, ospec in (("bf16 oracle", dict(s32, precision="bf16")), ("fp32 oracle", s32)):
        tight = yard == "bf16 oracle"
        fwd, loss, grads, otouched = O.train_step(ospec, data, y, err, 0.01 * scp, missing_mode=missing_mode, dropout_seed=seed)
        assert (touched == otouched).all()
        want = flat_grads(grads)
        if tight:
            assert_close(got, want, rtol=2e-3, what=f"{name}: grads vs {yard}")
        w64 = np.asarray(want, dtype=np.float64)
        cos = float(got @ w64 / (np.linalg.norm(got) * np.linalg.norm(w64)))
        assert cos >= (0.9999 if tight else 0.99), f"{name}: gradient direction vs {yard}: cos = {cos}"
        acc = O.EpochAccumulator(len(feats), D)
        acc.add(fwd)
        fin = acc.finalize()
        assert_close(hist.loss["train"][0], fin["loss"], rtol=1e-4 if tight else 1e-2, what=f"{name}: train loss vs {yard}")
        assert_close(hist.state_change_loss[0], fin["state_change"], rtol=2e-3 if tight else 1e-2, what=f"{name}: state change vs {yard}")
        for n in HIST[1:]:          # counters: a near-tie arg-max may flip
            np.testing.assert_allclose(np.nan_to_num(getattr(hist, n)["train"][0]), np.nan_to_num(fin[n]),
                                       atol=0.01 if tight else 0.03, err_msg=f"{n} vs {yard}")
        ofwd = O.forward(ospec, data, y, None, missing_mode)
        assert (pred != ofwd["predictions"]).mean() <= (0.005 if tight else 0.03), yard
        assert_close(states, ofwd["final_state"], rtol=8e-3 if tight else 1e-2, what=f"{name}: states vs {yard}")
        acc = O.EpochAccumulator(len(feats), D)
        acc.add(ofwd)
        assert_close(hist.loss["val"][0], acc.finalize()["loss"], rtol=1e-4 if tight else 1e-2, what=f"{name}: val loss vs {yard}")
    # missing rows keep their state bit for bit: a row with every modality absent ends on the bf16 image of s_0
    if mnar and missing_mode == "row":
        absent = np.all([np.isnan(x).any(1) for x in data], axis=0)
        if absent.any():
            s0 = torch.tensor(spec["init_state"], dtype=torch.float32).to(torch.bfloat16).float().numpy().reshape(-1)
            assert (states[absent] == s0[None, :]).all()
    return model
