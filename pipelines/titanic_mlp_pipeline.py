#!/usr/bin/env python
"""Titanic MLP pipeline on the B200 step (reference: pipelines/titanic/titanic_mlp_pipeline.py:19-127).

Same experiment, same hyper-parameters, same artefacts (model state, history pickle, results CSV);
the only differences are where the table comes from (``--csv``; ``--synthetic N`` writes a
Titanic-shaped table first because the real file cannot be downloaded here) and that the model is the
fused CUDA implementation.

    python pipelines/titanic_mlp_pipeline.py --synthetic 891 --epoch 30 --out-dir /tmp/titanic
"""
import argparse
import os
import pickle
import sys

import torch
import torch.nn.functional as F
from torch.nn import CrossEntropyLoss
from torch.utils.data import DataLoader

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from multimodn_b200 import MultiModN, MultiModNHistory  # noqa: E402
from multimodn_b200.datasets.titanic import TitanicDataset, write_synthetic_titanic_csv  # noqa: E402
from multimodn_b200.decoders import LogisticDecoder  # noqa: E402
from multimodn_b200.encoders import MLPEncoder  # noqa: E402

PIPELINE_NAME = "titanic_mlp_pipeline"
FEATURES = ["Fare", "Pclass", "Age", "Sex_male", "Relatives", "Embarked"]      # titanic_mlp_pipeline.py:26
TARGETS = ["Survived"]


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Titanic MLP pipeline for MultiModN on B200")
    p.add_argument("-e", "--epoch", type=int, default=300)
    p.add_argument("-s", "--seed", type=int, default=0)
    p.add_argument("--csv", help="table with the Kaggle Titanic schema")
    p.add_argument("--synthetic", type=int, metavar="N", help="write an N-row synthetic table to --out-dir and use it")
    p.add_argument("--out-dir", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "out"))
    p.add_argument("--device", default="cuda")
    p.add_argument("--no-save", action="store_true", help="skip the model / history / results files")
    return p.parse_args(argv)


def run(csv_path, epochs, seed, device, batch_size=32, on_model=None):
    """The experiment of titanic_mlp_pipeline.py:24-85; returns (model, history, (train, val) subsets).
    on_model(model), if given, runs after construction and before the first epoch."""
    torch.manual_seed(seed)
    datasplit = (0.8, 0.2, 0)
    state_size = 1
    learning_rate = 0.01
    dataset = TitanicDataset(FEATURES, TARGETS, csv_path, dropna=True, std=True).partition_dataset()
    train_data, val_data, _ = dataset.random_split(datasplit, seed, 0)       # balanced on 'Survived'
    train_loader = DataLoader(train_data, batch_size if batch_size else len(train_data))
    val_loader = DataLoader(val_data, batch_size if batch_size else len(val_data))

    encoders = [MLPEncoder(state_size, len(FEATURES), (5, 5), F.relu)]
    decoders = [LogisticDecoder(state_size) for _ in TARGETS]
    model = MultiModN(state_size, encoders, decoders, 0.7, 0.3, device=torch.device(device))
    if on_model is not None:
        on_model(model)
    optimizer = torch.optim.Adam(list(model.parameters()), learning_rate)
    criterion = CrossEntropyLoss()
    history = MultiModNHistory(TARGETS)
    for _ in range(epochs):
        model.train_epoch(train_loader, optimizer, criterion, history)
        model.test(val_loader, criterion, history, tag="val")
    return model, history, (train_data, val_data)


def main(argv=None):
    args = parse_args(argv)
    os.makedirs(args.out_dir, exist_ok=True)
    csv_path = args.csv
    if args.synthetic:
        csv_path = os.path.join(args.out_dir, "titanic_synthetic.csv")
        write_synthetic_titanic_csv(csv_path, args.synthetic, args.seed)
    if not csv_path:
        raise SystemExit("give --csv PATH or --synthetic N")
    model, history, _ = run(csv_path, args.epoch, args.seed, args.device)
    history.print_results()
    if not args.no_save:
        torch.save(model.state_dict(), os.path.join(args.out_dir, PIPELINE_NAME + "_model.pt"))
        with open(os.path.join(args.out_dir, PIPELINE_NAME + "_history.pkl"), "wb") as f:
            pickle.dump(history, f)
        history.save_results(os.path.join(args.out_dir, PIPELINE_NAME + ".csv"))
    return model, history


if __name__ == "__main__":
    main()
