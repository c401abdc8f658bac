"""the Titanic MLP pipeline end to end — table -> TitanicDataset -> stratified split ->
DataLoader -> MultiModN.train_epoch / test -> MultiModNHistory.get_results — against the same script
run with the UNMODIFIED reference classes (tests/golden/make_golden.py:fixture_titanic_pipeline)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import load_golden, golden_spec, flat_params, assert_close
from model_utils import model_spec
from multimodn_b200 import MultiModN
from multimodn_b200.datasets import TitanicDataset, write_synthetic_titanic_csv
from multimodn_b200.decoders import LogisticDecoder
from multimodn_b200.encoders import MLPEncoder

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pipelines import titanic_mlp_pipeline as pipeline  # noqa: E402

HIST = ("loss", "accuracy", "sensitivity", "specificity", "balanced_accuracy")


@pytest.fixture
def table(tmp_path):
    fx = load_golden("titanic_pipeline")
    csv = str(tmp_path / "titanic.csv")
    write_synthetic_titanic_csv(csv, int(fx["n_rows"]), int(fx["seed"]))
    return fx, csv


def test_dataset_and_split_match_reference(table):
    fx, csv = table
    ds = TitanicDataset(pipeline.FEATURES, pipeline.TARGETS, csv, dropna=True, std=True)
    assert ds.X.shape == fx["X"].shape
    np.testing.assert_allclose(ds.X, fx["X"], rtol=1e-12, atol=1e-12)
    assert (ds.y == fx["y"]).all()
    train, val, test = ds.partition_dataset().random_split((0.8, 0.2, 0), int(fx["seed"]), 0)
    assert (np.asarray(train.indices) == fx["train_idx"]).all()
    assert (np.asarray(val.indices) == fx["val_idx"]).all()
    assert len(test) == len(fx["test_idx"]) == 0
    data, target = train[3]
    assert isinstance(data, list) and data[0].dtype == torch.float32 and data[0].shape == (6,)
    assert target.shape == (1,)


def test_seeded_construction_matches_reference(table):
    """same seed -> bit-identical initial weights (module construction order of the pipeline)"""
    fx, _ = table
    torch.manual_seed(int(fx["seed"]))
    encoders = [MLPEncoder(1, 6, (5, 5), F.relu)]
    decoders = [LogisticDecoder(1)]
    model = MultiModN.__new__(MultiModN)          # construct the modules only: no device needed for this check
    torch.nn.Module.__init__(model)
    from multimodn_b200.state import TrainableInitState
    model.encoders, model.decoders = torch.nn.ModuleList(encoders), torch.nn.ModuleList(decoders)
    model.init_state = TrainableInitState(1, torch.device("cpu"))
    got, want = flat_params(model_spec(model)), flat_params(golden_spec(fx))
    assert (got == want).all()


def check_pipeline(fx, csv, device, on_model=None):
    model, history, (train, val) = pipeline.run(csv, int(fx["epochs"]), int(fx["seed"]), device,
                                                batch_size=int(fx["batch_size"]), on_model=on_model)
    assert len(train) == len(fx["train_idx"]) and len(val) == len(fx["val_idx"])
    for ep in range(int(fx["epochs"])):
        for name in HIST:
            assert_close(getattr(history, name)["train"][ep], fx[f"train_{name}"][ep], rtol=1e-5, what=f"train {name} {ep}")
            assert_close(getattr(history, name)["val"][ep], fx[f"val_{name}"][ep], rtol=1e-5, what=f"val {name} {ep}")
        assert_close(history.state_change_loss[ep], fx["train_state_change"][ep], rtol=1e-5, what=f"state change {ep}")
    assert_close(flat_params(model_spec(model)), flat_params(golden_spec(fx, "spec_final")), rtol=1e-5, what="weights")
    results = history.get_results()
    assert list(results.columns) == [str(c) for c in fx["results_columns"]]
    assert list(results.index) == [str(c) for c in fx["results_index"]]
    assert_close(results.to_numpy(dtype=np.float64), fx["results"], rtol=1e-5, what="results table")
    return model, history


def test_pipeline_matches_reference_emulated(emu, table):
    fx, csv = table
    check_pipeline(fx, csv, "cpu")


def test_pipeline_artifacts(emu, tmp_path):
    """the driver's files: results CSV, pickled history, state_dict that loads into a fresh model"""
    out = str(tmp_path / "out")
    model, history = pipeline.main(["--synthetic", "120", "--epoch", "1", "--out-dir", out, "--device", "cpu"])
    import pandas as pd
    import pickle
    df = pd.read_csv(os.path.join(out, "titanic_mlp_pipeline.csv"), index_col=0)
    assert list(df.index) == ["Survived"] and "Val balanced accuracy" in df.columns
    with open(os.path.join(out, "titanic_mlp_pipeline_history.pkl"), "rb") as f:
        again = pickle.load(f)
    assert np.array_equal(again.loss["train"][0], history.loss["train"][0])
    state = torch.load(os.path.join(out, "titanic_mlp_pipeline_model.pt"))
    fresh = MultiModN(1, [MLPEncoder(1, 6, (5, 5), F.relu)], [LogisticDecoder(1)], 0.7, 0.3, device=torch.device("cpu"))
    fresh.load_state_dict(state)
    assert (flat_params(model_spec(fresh)) == flat_params(model_spec(model))).all()


@pytest.mark.gpu
def test_pipeline_matches_reference_gpu(table):
    """On a CUDA device TrainableInitState draws from the CUDA generator (as the reference's does, state.py:25-27), so the
    seeded initial state differs from the CPU-generated golden run: start from the golden's initial weights instead."""
    fx, csv = table
    spec0 = golden_spec(fx)

    def load_golden_weights(model):
        with torch.no_grad():
            model.init_state.state_value.copy_(torch.from_numpy(np.asarray(spec0["init_state"], dtype=np.float32)).reshape(1, -1))
            lins = list(model.encoders[0].layers) + [model.decoders[0].fc]
            pairs = list(spec0["encoders"][0]["layers"]) + list(spec0["decoders"][0]["layers"])
            for lin, (W, b) in zip(lins, pairs):
                lin.weight.copy_(torch.from_numpy(np.asarray(W, dtype=np.float32)))
                lin.bias.copy_(torch.from_numpy(np.asarray(b, dtype=np.float32)))

    check_pipeline(fx, csv, "cuda", on_model=load_golden_weights)
