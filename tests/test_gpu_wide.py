"""GPU: the wide regime (precision="bf16": every layer a tcgen05 bf16 GEMM over the whole batch, mmn_wide*.cuh).

Two yardsticks:
  * the oracle's bf16 restatement (oracle.multimodn_oracle with spec["precision"] = "bf16": the same algorithm rounded to
    bfloat16 exactly where the CUDA path stores bfloat16) -- gradients, losses, states agree to summation-order accuracy;
  * the fp32 oracle (= the reference) -- per-step losses, history metrics and predictions within the north star's
    bf16 tolerance of 1e-2 relative.  Gradients are not compared at 1e-2 against fp32: rounding the layer gradients to
    bfloat16 moves individual weight-gradient tensors by several percent in ANY bf16 implementation (the bf16 oracle
    shows the same per-tensor deviations digit for digit, profiles/wide_grad_errors.py); their direction is checked."""
import numpy as np
import pytest
import torch
from torch.nn import CrossEntropyLoss

from oracle import multimodn_oracle as O
from oracle.spec_io import random_spec, synthetic_batch
from multimodn_b200 import MultiModNHistory, FusedAdam
from helpers import flat_grads, assert_close
from model_utils import model_from_spec, GradTap, tapped_flat

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def force_layerwise_regime(monkeypatch):
    """narrow models under precision="bf16" default to the per-tile kernel (tests/test_gpu_nb.py); this file is about the
    layer-wise tcgen05 regime, which MMN_ENGINE=wide selects for any model"""
    monkeypatch.setenv("MMN_ENGINE", "wide")
HIST = ("loss", "accuracy", "sensitivity", "specificity", "balanced_accuracy")
RTOL = 1e-2

# name: (S, features, enc_kind, enc_hidden, D, dec_hidden, n_classes, B, mnar, dropout)
CASES = {
    "ragged_dims": (60, [37, 70], "mimic", (90, 77), 2, (48,), 2, 300, True, 0.0),
    "mlp_kind": (40, [33, 18, 9], "mlp", (64, 24), 2, (), 2, 517, True, 0.0),
    "multiclass": (64, [128], "mimic", (128,), 3, (64, 32), 5, 256, False, 0.0),
    "dropout": (48, [64, 40], "mimic", (96, 96), 2, (32,), 2, 400, True, 0.3),
    "config4_shape": (1024, [1024, 768], "mimic", (2048, 2048), 2, (2048,), 2, 256, False, 0.0),
}


def run_case(name, missing_mode="row"):
    S, feats, kind, eh, D, dh, C, B, mnar, p = CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    spec = random_spec(rng, S, feats, enc_kind=kind, enc_hidden=eh, dropout=p, n_decoders=D, dec_hidden=dh, n_classes=C)
    data, y = synthetic_batch(rng, feats, D, B, mnar=mnar, n_classes=C)
    err, scp = 0.8, 0.6
    model = model_from_spec(spec, err, scp, DEV, missing_mode, precision="bf16")
    from multimodn_b200 import _lib
    assert _lib.get_lib().dll.mmn_plan_engine(model.runtime().plan) == 3
    tap = GradTap(model.parameters())
    hist = MultiModNHistory([str(i) for i in range(D)])
    loader = [([torch.from_numpy(x).to(DEV) for x in data], torch.from_numpy(y).to(DEV))]
    rt = model.runtime()
    rt.dropout_base_seed, rt.step_counter = 77, 0
    seed = (77 * 0x9E3779B1 + 1 * 0x85EBCA77) & 0xFFFFFFFF
    model.train_epoch(loader, tap, CrossEntropyLoss(), hist)
    got, touched = tapped_flat(model, tap)
    assert np.isfinite(got).all()
    pred = model.predict([torch.from_numpy(x) for x in data])
    states = torch.stack(model.get_states(loader)).cpu().numpy()

    s32 = O.cast_spec(spec, np.float32)
    for yard, ospec in (("bf16 oracle", dict(s32, precision="bf16")), ("fp32 oracle", s32)):
        tight = yard == "bf16 oracle"
        fwd, loss, grads, otouched = O.train_step(ospec, data, y, err, 0.01 * scp, missing_mode=missing_mode, dropout_seed=seed)
        assert (touched == otouched).all()
        want = flat_grads(grads).astype(np.float64)
        if tight:
            assert_close(got, want, rtol=2e-3, what=f"{name}: grads vs {yard}")
        cos = float(got @ want / (np.linalg.norm(got) * np.linalg.norm(want)))
        assert cos >= (0.9999 if tight else 0.99), f"{name}: gradient direction vs {yard}: cos = {cos}"
        acc = O.EpochAccumulator(len(feats), D)
        acc.add(fwd)
        fin = acc.finalize()
        assert_close(hist.loss["train"][0], fin["loss"], rtol=1e-4 if tight else RTOL, what=f"{name}: train loss vs {yard}")
        assert_close(hist.state_change_loss[0], fin["state_change"], rtol=2e-3 if tight else RTOL, what=f"{name}: state change vs {yard}")
        for n in HIST[1:]:          # counters: a near-tie arg-max may flip
            np.testing.assert_allclose(np.nan_to_num(getattr(hist, n)["train"][0]), np.nan_to_num(fin[n]),
                                       atol=0.01 if tight else 0.03, err_msg=f"{n} vs {yard}")
        ofwd = O.forward(ospec, data, y, None, missing_mode)
        assert (pred != ofwd["predictions"]).mean() <= (0.005 if tight else 0.03), yard
        assert_close(states, ofwd["final_state"], rtol=8e-3 if tight else RTOL, what=f"{name}: states vs {yard}")
    # missing rows keep their state bit for bit: a row with every modality absent ends on the bf16 image of s_0
    if mnar:
        absent = np.all([np.isnan(x).any(1) for x in data], axis=0)
        if absent.any():
            s0 = torch.tensor(spec["init_state"], dtype=torch.float32).to(torch.bfloat16).float().numpy().reshape(-1)
            assert (states[absent] == s0[None, :]).all()
    return model


@pytest.mark.parametrize("name", sorted(CASES))
def test_wide_case_matches_oracle(name):
    run_case(name)


def test_wide_batch_missing_mode():
    run_case("ragged_dims", missing_mode="batch")


def test_wide_training_reduces_loss_and_matches_fp32_path():
    """a few Adam steps: the bf16 regime follows the fp32 fused kernels' trajectory to bf16 accuracy"""
    S, feats = 64, [48, 80]
    rng = np.random.default_rng(5)
    spec = random_spec(rng, S, feats, enc_kind="mimic", enc_hidden=(64, 64), dropout=0.0, n_decoders=2, dec_hidden=(32,), n_classes=2)
    data, y = synthetic_batch(rng, feats, 2, 1024, mnar=True)
    losses = {}
    for prec in ("fp32", "bf16"):
        model = model_from_spec(spec, 1.0, 0.3, DEV, "row", precision=prec)
        opt = FusedAdam(model, lr=1e-2)
        hist = MultiModNHistory(["a", "b"])
        loader = [([torch.from_numpy(x).to(DEV) for x in data], torch.from_numpy(y).to(DEV))]
        for _ in range(8):
            model.train_epoch(loader, opt, CrossEntropyLoss(), hist)
        losses[prec] = np.array([m[-1].sum() for m in hist.loss["train"]])
    assert losses["bf16"][-1] < losses["bf16"][0]
    np.testing.assert_allclose(losses["bf16"], losses["fp32"], rtol=2e-2)


def test_fp32_plan_rejects_wide_model_with_pointer_to_bf16():
    from multimodn_b200 import _lib
    rng = np.random.default_rng(0)
    spec = random_spec(rng, 1024, [64], enc_kind="mimic", enc_hidden=(2048,), dropout=0.0, n_decoders=1, dec_hidden=(), n_classes=2)
    model = model_from_spec(spec, 1.0, 0.3, DEV, "row")
    with pytest.raises(_lib.MMNError, match="precision = bf16"):
        model.runtime()


@pytest.mark.parametrize("B", [1, 7, 65])
def test_wide_tiny_batches(B):
    """fewer rows than one TMA box / MMA tile / 64 x 64 element-wise tile: everything out of range is zero-filled"""
    S, feats = 40, [33, 18]
    rng = np.random.default_rng(B)
    spec = random_spec(rng, S, feats, enc_kind="mimic", enc_hidden=(64, 24), dropout=0.0, n_decoders=2, dec_hidden=(16,), n_classes=2)
    data, y = synthetic_batch(rng, feats, 2, B, mnar=False)
    model = model_from_spec(spec, 0.8, 0.6, DEV, "row", precision="bf16")
    tap = GradTap(model.parameters())
    hist = MultiModNHistory(["a", "b"])
    loader = [([torch.from_numpy(x).to(DEV) for x in data], torch.from_numpy(y).to(DEV))]
    model.train_epoch(loader, tap, CrossEntropyLoss(), hist)
    got, _ = tapped_flat(model, tap)
    ospec = dict(O.cast_spec(spec, np.float32), precision="bf16")
    fwd, loss, grads, _ = O.train_step(ospec, data, y, 0.8, 0.006)
    assert_close(got, flat_grads(grads), rtol=2e-3, what="grads")
    assert_close(hist.loss["train"][0], fwd["ce"], rtol=1e-4, what="loss")
    pred = model.predict([torch.from_numpy(x) for x in data])
    assert pred.shape == (3, 2, B)
    assert (pred != O.forward(ospec, data, y)["predictions"]).mean() <= 0.15 if B > 1 else True


def test_wide_shard_additivity():
    """the data-parallel contract in the wide regime: two row shards normalised by the GLOBAL batch add up to the full
    batch (gradients, losses, present counts); the dropout stream is keyed by the global row index"""
    S, feats, B = 48, [64, 40], 512
    rng = np.random.default_rng(11)
    spec = random_spec(rng, S, feats, enc_kind="mimic", enc_hidden=(96, 96), dropout=0.25, n_decoders=2, dec_hidden=(32,), n_classes=2)
    data, y = synthetic_batch(rng, feats, 2, B, mnar=True)
    model = model_from_spec(spec, 1.0, 0.3, DEV, "row", precision="bf16")
    rt = model.runtime()
    dev = [torch.from_numpy(x).to(DEV) for x in data]
    ty = torch.from_numpy(y).to(DEV)
    seq = [(i, i) for i in range(len(feats))]

    def run(shards):
        acc_g, acc_m = torch.zeros_like(rt.gflat), rt.new_metrics()
        for r in range(shards):
            n = B // shards
            rt.step_counter = 0                      # same dropout seed for every call
            mb, keep, rows = rt.prepare_batch([t[r * n:(r + 1) * n] for t in dev], ty[r * n:(r + 1) * n], seq, "row",
                                              (shards, r, None))
            rt.train_step(mb, rows, 1.0, 0.003, True, acc_m)
            acc_g += rt.gflat
        return acc_g.cpu().numpy(), acc_m.cpu().numpy()

    g1, m1 = run(1)
    g2, m2 = run(2)
    assert_close(g2[:rt.packed.n_params], g1[:rt.packed.n_params], rtol=1e-4, what="sharded grads == full-batch grads")
    assert (g2[rt.packed.n_params:] == g1[rt.packed.n_params:]).all()          # present-row counts
    assert_close(m2, m1, rtol=1e-5, what="sharded metrics == full-batch metrics")


def test_wide_test_epoch_history_and_outputs():
    """MultiModN.test() on a bf16 plan: the 'val' history row and the collected last-step outputs follow the oracle"""
    S, feats, B = 48, [64, 40], 384
    rng = np.random.default_rng(21)
    spec = random_spec(rng, S, feats, enc_kind="mimic", enc_hidden=(96, 64), dropout=0.0, n_decoders=2, dec_hidden=(32,), n_classes=2)
    data, y = synthetic_batch(rng, feats, 2, B, mnar=True)
    model = model_from_spec(spec, 1.0, 0.3, DEV, "row", precision="bf16")
    hist = MultiModNHistory(["a", "b"])
    loader = [([torch.from_numpy(x[i:i + 128]).to(DEV) for x in data], torch.from_numpy(y[i:i + 128]).to(DEV)) for i in range(0, B, 128)]
    results = model.test(loader, CrossEntropyLoss(), hist, tag="val")
    assert len(results) == 2
    ospec = dict(O.cast_spec(spec, np.float32), precision="bf16")
    acc = O.EpochAccumulator(len(feats), 2)
    for i in range(0, B, 128):
        acc.add(O.forward(ospec, [x[i:i + 128] for x in data], y[i:i + 128], None, "row"))
    fin = acc.finalize()
    assert_close(hist.loss["val"][0], fin["loss"], rtol=1e-4, what="val loss")
    np.testing.assert_allclose(np.nan_to_num(hist.accuracy["val"][0]), np.nan_to_num(fin["accuracy"]), atol=0.01)


def test_wide_full_config4_at_8192_rows():
    """BASELINE configs[3] complete (4 encoders 1024/1024/768/768 -> 2048 -> 2048 -> state 1024, 2 decoders 1024 -> 2048 -> 2,
    MNAR missingness) at its per-GPU batch of 8192 rows.  The oracle cannot run 8192 rows of this model in test time, so:
      * per-row results (predictions after every step, final states) of the 8192-row calls are compared with the bf16 oracle
        on a 512-row sample of the same batch -- rows are independent in the forward pass;
      * the 8192-row gradient / metric vector equals the sum of its sixteen 512-row shards normalised by the global batch
        (the data-parallel contract), and the first shard alone, run through train_epoch, agrees with the bf16 oracle's
        gradient of those 512 rows."""
    from oracle.spec_io import config_spec, CONFIGS
    feats, D, B, n = CONFIGS["c4_wide"]["features"], CONFIGS["c4_wide"]["n_decoders"], 8192, 512
    spec = config_spec("c4_wide", 5)
    rng = np.random.default_rng(2024)
    data, y = synthetic_batch(rng, feats, D, B, mnar=True)
    err, scp = 0.8, 0.6
    model = model_from_spec(spec, err, scp, DEV, "row", precision="bf16")
    rt = model.runtime()
    from multimodn_b200 import _lib
    assert _lib.get_lib().dll.mmn_plan_engine(rt.plan) == 3
    ospec = dict(O.cast_spec(spec, np.float32), precision="bf16")

    # forward, per row
    pred = model.predict([torch.from_numpy(x) for x in data])
    loader = [([torch.from_numpy(x).to(DEV) for x in data], torch.from_numpy(y).to(DEV))]
    states = torch.stack(model.get_states(loader)).cpu().numpy().reshape(B, -1)
    idx = np.sort(rng.choice(B, n, replace=False))
    ofwd = O.forward(ospec, [x[idx] for x in data], y[idx], None, "row")
    assert pred.shape == (len(feats) + 1, D, B)
    assert (pred[:, :, idx] != ofwd["predictions"]).mean() <= 0.005
    assert_close(states[idx], ofwd["final_state"].reshape(n, -1), rtol=8e-3, what="states of the sampled rows vs bf16 oracle")

    # backward: the full batch == the sum of its shards; one shard == the oracle
    dev = [torch.from_numpy(x).to(DEV) for x in data]
    ty = torch.from_numpy(y).to(DEV)
    seq = [(i, i) for i in range(len(feats))]

    def run(shards):
        acc_g, acc_m = torch.zeros_like(rt.gflat), rt.new_metrics()
        for r in range(shards):
            m = B // shards
            rt.step_counter = 0
            mb, keep, rows = rt.prepare_batch([t[r * m:(r + 1) * m] for t in dev], ty[r * m:(r + 1) * m], seq, "row", (shards, r, None))
            rt.train_step(mb, rows, err, 0.01 * scp, True, acc_m)
            acc_g += rt.gflat
        return acc_g.cpu().numpy(), acc_m.cpu().numpy()

    g1, m1 = run(1)
    g16, m16 = run(B // n)
    P = rt.packed.n_params
    assert np.isfinite(g1).all()
    assert_close(g16[:P], g1[:P], rtol=2e-3, what="sum of 16 shards == the 8192-row step (gradients)")
    assert (g16[P:] == g1[P:]).all()
    assert_close(m16, m1, rtol=1e-5, what="sum of 16 shards == the 8192-row step (metrics)")
    # one 512-row shard through the public API (normalised by its own size) against the oracle
    sl = slice(0, n)
    tap = GradTap(model.parameters())
    model.train_epoch([([t[sl] for t in dev], ty[sl])], tap, CrossEntropyLoss(), MultiModNHistory(["a", "b"]))
    got, _ = tapped_flat(model, tap)
    _, _, grads, _ = O.train_step(ospec, [x[sl] for x in data], y[sl], err, 0.01 * scp, missing_mode="row")
    want = flat_grads(grads).astype(np.float64)
    cos = float(got @ want / (np.linalg.norm(got) * np.linalg.norm(want)))
    assert cos >= 0.9999, cos
    assert_close(got, want, rtol=2e-3, what="first 512-row shard vs bf16 oracle (gradients)")
