"""GPU (B200): the product path through libmmn.so (C ABI, sm_100a kernels) against
 (a) the golden vectors produced by the unmodified reference (tests/golden), and
 (b) the oracle on seeded synthetic inputs (tests/parity_cases.py), and
 (c) size-independent properties at BASELINE.json's full sizes.
Tolerance: 1e-5 relative (norm-wise) in fp32, as BASELINE.json's north_star states; predictions,
counters and missing-row pass-through are exact."""
import numpy as np
import pytest
import torch
from torch.nn import CrossEntropyLoss

pytestmark = pytest.mark.gpu

from multimodn_b200 import MultiModNHistory, FusedAdam  # noqa: E402
from helpers import load_golden, golden_data, golden_spec, golden_grads, flat_grads, flat_params, assert_close  # noqa: E402
from model_utils import model_from_spec, model_spec, GradTap, tapped_flat, batches  # noqa: E402
from parity_cases import CASES, run_parity_case  # noqa: E402
import test_emu_golden as G  # noqa: E402  (shares the case runner; it is device-agnostic)

DEV = "cuda"


@pytest.fixture(autouse=True)
def _need_cuda():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from multimodn_b200 import _lib
    lib = _lib.get_lib()            # raises if libmmn.so is missing: no fallback
    assert lib.host_memory is False


def run_golden(name, names, missing_mode, **kw):
    fx = load_golden(name)
    spec = golden_spec(fx)
    data, y = golden_data(fx), fx["y"]
    seq = kw.pop("seq", None)
    bs = kw.pop("bs", None) or len(y)
    model = model_from_spec(spec, float(fx["err_penalty"]), float(fx["state_change_penalty"]), DEV, missing_mode)
    tap = GradTap(model.parameters())
    hist = MultiModNHistory(names)
    model.train_epoch(batches(data, y, bs, seq, DEV), tap, CrossEntropyLoss(), hist)
    got, touched = tapped_flat(model, tap)
    assert_close(got, flat_grads(golden_grads(fx, spec)), rtol=1e-5, what="grads")
    G.check_history(fx, "train", hist, "train", with_sc=True)
    if kw.get("check_val", True):
        model.test(batches(data, y, bs, seq, DEV), CrossEntropyLoss(), hist, tag="val")
        G.check_history(fx, "val", hist, "val")
    if kw.get("check_predict", True) and "predict" in fx:
        assert (model.predict([torch.from_numpy(x) for x in data], seq) == fx["predict"]).all()
    states = torch.stack(model.get_states(batches(data, y, bs, seq, DEV))).cpu().numpy()
    assert_close(states, fx["states"], rtol=1e-5, what="states")
    return model, touched, fx


def test_golden_c2_small():
    run_golden("c2_mimic_small", ["a", "b"], "batch")


def test_golden_c2_full_dims():
    run_golden("c2_mimic_full", ["a", "b"], "row")


def test_golden_sequence():
    fx = load_golden("sequence")
    run_golden("sequence", ["a", "b", "c"], "row", seq=fx["seq"])


def test_golden_zoo():
    run_golden("zoo", ["a", "b", "c"], "row", check_val=False)


def test_golden_missing_row():
    model, touched, fx = run_golden("missing_row", ["a", "b"], "row", check_predict=False)
    assert (touched == fx["touched"]).all() and not touched[1]


def test_golden_missing_batch():
    fx = load_golden("missing_batch")
    spec = golden_spec(fx)
    data, y, B = golden_data(fx), fx["y"], int(fx["batch_size"])
    model = model_from_spec(spec, 1.0, 0.5, DEV, "batch")
    tap = GradTap(model.parameters())
    hist = MultiModNHistory(["a", "b"])
    model.train_epoch(batches(data, y, B, device=DEV)[:1], tap, CrossEntropyLoss(), hist)
    got, touched = tapped_flat(model, tap)
    assert list(touched) == [True, False, True]
    assert_close(got, flat_grads(golden_grads(fx, spec)), rtol=1e-5, what="grads")
    G.check_history(fx, "train1", hist, "train", with_sc=True)
    hist2 = MultiModNHistory(["a", "b"])
    model.train_epoch(batches(data, y, B, device=DEV), GradTap(model.parameters()), CrossEntropyLoss(), hist2)
    model.test(batches(data, y, B, device=DEV), CrossEntropyLoss(), hist2, tag="val")
    G.check_history(fx, "train2", hist2, "train", with_sc=True)
    G.check_history(fx, "val2", hist2, "val")


@pytest.mark.parametrize("opt_kind", ["torch_adam", "fused_adam"])
def test_golden_c1_titanic_two_epochs(opt_kind):
    fx = load_golden("c1_titanic")
    spec = golden_spec(fx)
    x, y, vx, vy = [fx["x0"]], fx["y"], [fx["vx0"]], fx["vy"]
    bs = int(fx["batch_size"])
    model = model_from_spec(spec, 0.7, 0.3, DEV, "row")
    if opt_kind == "torch_adam":
        opt = torch.optim.Adam(list(model.parameters()), float(fx["lr"]))
    else:
        opt = FusedAdam(model, lr=float(fx["lr"]))
    hist = MultiModNHistory(["Survived"])
    for ep in range(2):
        model.train_epoch(batches(x, y, bs, device=DEV), opt, CrossEntropyLoss(), hist)
        model.test(batches(vx, vy, bs, device=DEV), CrossEntropyLoss(), hist, tag="val")
    for ep in range(2):
        G.check_history(fx, "train", hist, "train", idx=ep, with_sc=True)
        G.check_history(fx, "val", hist, "val", idx=ep)
    assert_close(flat_params(model_spec(model)), flat_params(golden_spec(fx, "spec2")), rtol=1e-5, what="params")
    assert (model.predict([torch.from_numpy(vx[0])]) == fx["predict"]).all()


def test_golden_c2_adam_three_steps_fused():
    fx = load_golden("c2_mimic_small")
    spec = golden_spec(fx)
    data, y = golden_data(fx), fx["y"]
    model = model_from_spec(spec, 1.0, 0.3, DEV, "row")
    opt = FusedAdam(model, lr=float(fx["lr"]))
    for _ in range(3):
        model.train_epoch(batches(data, y, len(y), device=DEV), opt, CrossEntropyLoss())
    assert_close(flat_params(model_spec(model)), flat_params(golden_spec(fx, "spec3")), rtol=1e-5, what="params")


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_case(name):
    run_parity_case(name, DEV)


def test_batch_mode_oracle_case():
    run_parity_case("mnar_rows", DEV, missing_mode="batch")


# ---- full-size, size-independent properties (BASELINE.json configs 2, 3, 5) ----------------------
def _c3_model_and_data(B, seed=2, mnar=True):
    from oracle.spec_io import config_spec, synthetic_batch, CONFIGS
    spec = config_spec("c3_mnar", seed)
    rng = np.random.default_rng(seed)
    data, y = synthetic_batch(rng, CONFIGS["c3_mnar"]["features"], 6, B, mnar=mnar)
    model = model_from_spec(spec, 1.0, 0.3, DEV, "row")
    return spec, model, data, y


def test_c3_full_batch_shard_additivity_and_sample():
    """C3 at B=65536 with 30% MNAR: (1) a sample of rows agrees with the oracle; (2) metric sums
    and gradients are additive over row shards (the data-parallel contract): running two halves
    with n_rows_global = B reproduces the full-batch result."""
    B = 65536
    spec, model, data, y = _c3_model_and_data(B)
    dev = [torch.from_numpy(x).to(DEV) for x in data]
    ty = torch.from_numpy(y).to(DEV)
    tap = GradTap(model.parameters())
    hist = MultiModNHistory([str(i) for i in range(6)])
    model.train_epoch([(dev, ty)], tap, CrossEntropyLoss(), hist)
    full, _ = tapped_flat(model, tap)
    # fraction of missing cells is what the config promises
    miss = np.mean([np.isnan(x[:, 0]).mean() for x in data])
    assert 0.27 < miss < 0.33
    # oracle on a 512-row sample: states + predictions
    from oracle import multimodn_oracle as O
    idx = np.arange(0, B, B // 512)[:512]
    sub = [x[idx] for x in data]
    ofwd = O.forward(O.cast_spec(spec, np.float32), sub, y[idx], None, "row")
    states = torch.stack(model.get_states([(dev, ty)])).cpu().numpy()
    assert_close(states[idx], ofwd["final_state"], rtol=1e-5, what="c3 states sample")
    pred = model.predict(dev)
    assert (pred[:, :, idx] != ofwd["predictions"]).mean() < 2e-3
    # missing rows pass the state through bit-exactly: rows with every modality missing keep s_0
    all_missing = np.all([np.isnan(x[:, 0]) for x in data], axis=0)
    if all_missing.any():
        s0 = spec["init_state"].astype(np.float32)
        assert (states[all_missing] == s0[None, :]).all()
    # shard additivity through the C ABI (what the N-GPU path relies on)
    rt = model.runtime()
    import ctypes as C
    from multimodn_b200 import _lib as L
    acc_g = torch.zeros_like(rt.gflat)
    acc_m = rt.new_metrics()
    half = B // 2
    for r in range(2):
        shard = [t[r * half:(r + 1) * half] for t in dev]
        mb, keep, n = rt.prepare_batch(shard, ty[r * half:(r + 1) * half], [(i, i) for i in range(8)], "row",
                                       (2, r, None))
        rt.train_step(mb, n, 1.0, 0.003, True, acc_m)
        acc_g += rt.gflat
    tap2 = GradTap(model.parameters())
    rt.gflat.copy_(acc_g)
    rt.assign_grads()
    tap2.step()
    both, _ = tapped_flat(model, tap2)
    assert_close(both, full, rtol=2e-5, what="sharded grads == full-batch grads")
    mats, n_present, sc = rt.split_metrics(acc_m.cpu().numpy())
    assert_close(mats[0], hist.loss["train"][0], rtol=1e-6, what="sharded CE == full CE")
    assert n_present[0] == B


def test_c5_predict_permutation_properties():
    """C5: predict over permuted encoding sequences at N = 2^17 rows/launch: row 0 (initial state)
    never depends on the order; the row of the FIRST encoder of a sequence equals the row that
    encoder gets when it runs alone; the identity permutation equals encoder_sequence=None."""
    N = 1 << 17
    spec, model, data, y = _c3_model_and_data(N, seed=4, mnar=False)
    dev = [torch.from_numpy(x).to(DEV) for x in data]
    base = model.predict(dev)
    assert base.shape == (9, 6, N)
    ident = model.predict(dev, np.arange(8))
    assert (ident == base).all()
    rng = np.random.default_rng(4)
    for _ in range(3):
        perm = rng.permutation(8)
        # data is indexed by POSITION: give position i the modality the encoder perm[i] expects
        xs = [dev[e] for e in perm]
        out = model.predict(xs, perm)
        assert (out[0] == base[0]).all()
        first = int(perm[0])
        alone = model.predict([dev[first]], np.array([first]))
        assert (out[first + 1] == alone[first + 1]).all()


# ---- engines of fp32 plans: the FP32-FMA kernel trains; forward-only launches use the TMEM-resident tcgen05 kernel -----
def test_default_engine_is_fma():
    from multimodn_b200 import _lib
    fx = load_golden("c2_mimic_full")
    model = model_from_spec(golden_spec(fx), 1.0, 0.3, DEV, "row")
    assert _lib.get_lib().dll.mmn_plan_engine(model.runtime().plan) == 0


# ---- forward-only launches (test / predict / get_states) default to the TMEM-resident kernel where the model
#      qualifies; MMN_ENGINE=fma keeps them on the FP32-FMA kernel ---------------------------------------------------
def test_forward_engine_default_is_tmem_resident_for_c2():
    from multimodn_b200 import _lib
    fx = load_golden("c2_mimic_full")
    model = model_from_spec(golden_spec(fx), 1.0, 0.3, DEV, "row")
    assert _lib.get_lib().dll.mmn_plan_forward_engine(model.runtime().plan) == 2


@pytest.mark.parametrize("name,names,mode", [("c2_mimic_full", ["a", "b"], "row"), ("sequence", ["a", "b", "c"], "row"),
                                             ("zoo", ["a", "b", "c"], "row")])
def test_golden_fma_forward(monkeypatch, name, names, mode):
    monkeypatch.setenv("MMN_ENGINE", "fma")
    fx = load_golden(name)
    model, _, _ = run_golden(name, names, mode, seq=fx.get("seq"), check_val=name != "zoo")
    from multimodn_b200 import _lib
    assert _lib.get_lib().dll.mmn_plan_forward_engine(model.runtime().plan) == 0
