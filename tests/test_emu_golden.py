"""CPU: the product path (multimodn_b200.MultiModN -> C ABI -> the kernel SOURCES compiled for the
host by tests/emu) against the golden vectors of the unmodified reference.  This debugs the
kernels' indexing and the host bookkeeping without a GPU; tests/test_gpu_parity.py repeats the
same checks on the real library."""
import numpy as np
import pytest
import torch
from torch.nn import CrossEntropyLoss

from multimodn_b200 import MultiModNHistory
from helpers import load_golden, golden_data, golden_spec, golden_grads, flat_grads, flat_params, assert_close
from model_utils import model_from_spec, model_spec, GradTap, tapped_flat, batches

HIST = ("loss", "accuracy", "sensitivity", "specificity", "balanced_accuracy")


def check_history(fx, prefix, history, tag, idx=0, with_sc=False, rtol=1e-5):
    for name in HIST:
        assert_close(getattr(history, name)[tag][idx], fx[f"{prefix}_{name}"][idx], rtol=rtol, what=f"{prefix}_{name}")
    if with_sc:
        assert_close(history.state_change_loss[idx], fx[f"{prefix}_state_change"][idx], rtol=rtol, what="state_change")


def run_case(fx, names, missing_mode, seq=None, bs=None, grad_scale=1.0, decoders=2, rtol=1e-5, check_val=True,
             check_predict=True):
    spec = golden_spec(fx)
    data, y = golden_data(fx), fx["y"]
    bs = bs or len(y)
    model = model_from_spec(spec, float(fx["err_penalty"]), float(fx["state_change_penalty"]), "cpu", missing_mode)
    tap = GradTap(model.parameters())
    hist = MultiModNHistory(names)
    model.train_epoch(batches(data, y, bs, seq), tap, CrossEntropyLoss(), hist)
    got, touched = tapped_flat(model, tap, grad_scale)
    assert_close(got, flat_grads(golden_grads(fx, spec)), rtol=rtol, what="grads")
    check_history(fx, "train", hist, "train", with_sc=True, rtol=rtol)
    if check_val:
        model.test(batches(data, y, bs, seq), CrossEntropyLoss(), hist, tag="val")
        check_history(fx, "val", hist, "val", rtol=rtol)
    if check_predict and "predict" in fx:
        pred = model.predict([torch.from_numpy(x) for x in data], seq)
        assert pred.dtype == np.float64 and (pred == fx["predict"]).all()
    if "states" in fx:
        states = torch.stack(model.get_states(batches(data, y, bs, seq))).numpy()
        assert_close(states, fx["states"], rtol=rtol, what="states")
    return model, touched


def test_c2_small(emu):
    run_case(load_golden("c2_mimic_small"), ["a", "b"], "batch")


def test_c2_small_row_mode_equals_batch_mode_without_nan(emu):
    run_case(load_golden("c2_mimic_small"), ["a", "b"], "row")


def test_c2_full_dims(emu):
    run_case(load_golden("c2_mimic_full"), ["a", "b"], "row")


def test_permuted_sequence(emu):
    fx = load_golden("sequence")
    run_case(fx, ["a", "b", "c"], "row", seq=fx["seq"])


def test_module_zoo(emu):
    run_case(load_golden("zoo"), ["a", "b", "c"], "row", check_val=False)


def test_missing_row(emu):
    fx = load_golden("missing_row")
    model, touched = run_case(fx, ["a", "b"], "row", check_predict=False)
    assert (touched == fx["touched"]).all() and not touched[1]


def test_missing_batch(emu):
    fx = load_golden("missing_batch")
    spec = golden_spec(fx)
    data, y, B = golden_data(fx), fx["y"], int(fx["batch_size"])
    model = model_from_spec(spec, 1.0, 0.5, "cpu", "batch")
    tap = GradTap(model.parameters())
    hist = MultiModNHistory(["a", "b"])
    model.train_epoch(batches(data, y, B)[:1], tap, CrossEntropyLoss(), hist)
    got, touched = tapped_flat(model, tap)
    assert list(touched) == [True, False, True]
    assert_close(got, flat_grads(golden_grads(fx, spec)), rtol=1e-5, what="grads")
    check_history(fx, "train1", hist, "train", with_sc=True)
    hist2 = MultiModNHistory(["a", "b"])
    model.train_epoch(batches(data, y, B), GradTap(model.parameters()), CrossEntropyLoss(), hist2)
    model.test(batches(data, y, B), CrossEntropyLoss(), hist2, tag="val")
    check_history(fx, "train2", hist2, "train", with_sc=True)
    check_history(fx, "val2", hist2, "val")
    states = torch.stack(model.get_states(batches(data, y, B))).numpy()
    assert_close(states, fx["states"], rtol=1e-5, what="states")


def test_c1_titanic_two_epochs_adam(emu):
    fx = load_golden("c1_titanic")
    spec = golden_spec(fx)
    x, y, vx, vy = [fx["x0"]], fx["y"], [fx["vx0"]], fx["vy"]
    bs = int(fx["batch_size"])
    model = model_from_spec(spec, 0.7, 0.3, "cpu", "row")
    opt = torch.optim.Adam(list(model.parameters()), float(fx["lr"]))
    hist = MultiModNHistory(["Survived"])
    for ep in range(2):
        model.train_epoch(batches(x, y, bs), opt, CrossEntropyLoss(), hist)
        model.test(batches(vx, vy, bs), CrossEntropyLoss(), hist, tag="val")
    for ep in range(2):
        check_history(fx, "train", hist, "train", idx=ep, with_sc=True)
        check_history(fx, "val", hist, "val", idx=ep)
    assert_close(flat_params(model_spec(model)), flat_params(golden_spec(fx, "spec2")), rtol=1e-5, what="params")
    assert (model.predict([torch.from_numpy(vx[0])]) == fx["predict"]).all()


def test_dropout_fixed_mask(emu):
    fx = load_golden("dropout")
    spec = golden_spec(fx)
    data, y = golden_data(fx), fx["y"]
    model = model_from_spec(spec, 1.0, 0.3, "cpu", "row")
    rt = model.runtime()
    seed = int(fx["dropout_seed"])
    # pin the step's dropout stream key to the fixture's
    orig = rt.train_step

    def pinned(batch, n_rows, err, scp, training, metrics):
        rt.step_counter = 0
        rt.dropout_base_seed = 0
        import multimodn_b200._lib as L
        import ctypes as C
        rt.ensure_packed()
        ws, ws_bytes = rt.workspace(n_rows, True)
        targs = L.TrainArgs(err, scp, seed, 1)
        o = rt.outputs(metrics=metrics)
        rt.lib.check(rt.lib.dll.mmn_train_step(rt.plan, C.byref(batch), rt.flat.data_ptr(), C.byref(targs), C.byref(o),
                                               rt.gflat.data_ptr(), ws.data_ptr(), ws_bytes, rt.stream()))
    rt.train_step = pinned
    tap = GradTap(model.parameters())
    hist = MultiModNHistory(["a", "b"])
    model.train_epoch(batches(data, y, len(y)), tap, CrossEntropyLoss(), hist)
    got, _ = tapped_flat(model, tap)
    assert_close(got, flat_grads(golden_grads(fx, spec)), rtol=1e-5, what="grads")
    check_history(fx, "train", hist, "train", with_sc=True)
    rt.train_step = orig
    model.test(batches(data, y, len(y)), CrossEntropyLoss(), hist, tag="val")
    check_history(fx, "val", hist, "val")
