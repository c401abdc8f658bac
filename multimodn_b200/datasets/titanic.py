"""Titanic table -> modalities (reference: datasets/titanic/titanic_dataset.py).

``TitanicDataset`` reads a CSV with the Kaggle Titanic schema, derives the columns the
pipelines ask for (Relatives, Sex_male, Cabin_num, numeric Embarked), optionally drops rows
with missing values and standardises the features, and hands the matrix to
``PartitionDataset`` / ``FeatureWiseDataset``.  The reference resolves the file relative to its
own source tree (titanic_dataset.py:22); here the path is an argument.

``write_synthetic_titanic_csv`` produces a table of the same schema from a seeded generator
(no network in the build environment, so the real file cannot be fetched by get_data.sh).
"""
from itertools import accumulate
from typing import List, Optional, Sequence

import numpy as np
import pandas as pd
from torch import Tensor
from torch.utils.data import Dataset

from .multimod_dataset import FeatureWiseDataset, PartitionDataset

EMBARKED_CODES = {"S": 0, "C": 1, "Q": 2}          # titanic_dataset.py:79


def derive_columns(table: pd.DataFrame) -> pd.DataFrame:
    """Engineered columns of titanic_dataset.py:69-81."""
    out = table.copy()
    out["Relatives"] = out["SibSp"] + out["Parch"]
    out["Sex_male"] = (out["Sex"] == "male")          # get_dummies(..., drop_first=True): female is the dropped level
    out.loc[out["Sex"].isna(), "Sex_male"] = False
    out = out.drop(columns=["Sex"])
    cabins = sorted(out["Cabin"].dropna().unique())
    out["Cabin_num"] = out["Cabin"].map({name: i for i, name in enumerate(cabins)})
    out["Embarked"] = out["Embarked"].map(EMBARKED_CODES)
    return out


class TitanicDataset(Dataset):
    def __init__(self, features: List[str], targets: List[str], csv_path: str, dropna: bool = True,
                 dropna_columns: Sequence[str] = (), std: bool = True):
        table = pd.read_csv(csv_path).set_index("PassengerId")
        table["id"] = table.index
        table = derive_columns(table)
        wanted = list(dict.fromkeys(list(features) + list(targets) + list(dropna_columns)))
        table = table[wanted]
        if dropna:
            table = table.dropna()
        table = table[list(features) + list(targets)]
        X = table[list(features)].to_numpy(dtype=np.float64)
        if std:
            # StandardScaler: population standard deviation, constant columns divide by 1; NaNs ignored
            mean = np.nanmean(X, axis=0)
            scale = np.nanstd(X, axis=0)
            scale[scale == 0.0] = 1.0
            X = (X - mean) / scale
        self.X = X
        self.y = table[list(targets)].to_numpy()

    def __len__(self) -> int:
        return len(self.y)

    def __getitem__(self, idx: int):
        return Tensor(self.X[idx]), self.y[idx]

    def partition_dataset(self, partitions: Optional[List[int]] = None) -> PartitionDataset:
        return PartitionDataset(self.X, self.y, partitions)

    def featurewise_dataset(self) -> FeatureWiseDataset:
        return FeatureWiseDataset(self.X, self.y)

    def split_dataset(self, partitions: Optional[List[int]] = None) -> List[PartitionDataset]:
        if partitions is None:
            partitions = [self.X.shape[1]]
        if sum(partitions) != self.X.shape[1]:
            raise ValueError("Paritions sum doesn't match data dimension. Expected: {}, got: {}"
                             .format(sum(partitions), self.X.shape[1]))
        blocks = np.split(self.X, list(accumulate(partitions[:-1])), axis=1)
        return [PartitionDataset(block, self.y, [width]) for block, width in zip(blocks, partitions)]


def write_synthetic_titanic_csv(path: str, n_rows: int = 891, seed: int = 0) -> None:
    """A seeded table with the Kaggle schema and roughly its marginals: ~20 % of Age and ~77 % of
    Cabin missing, two missing Embarked values, survival correlated with sex / class / fare."""
    rng = np.random.default_rng(seed)
    pclass = rng.choice([1, 2, 3], size=n_rows, p=[0.24, 0.21, 0.55])
    male = rng.random(n_rows) < 0.65
    age = np.clip(rng.normal(29.7, 14.5, n_rows), 0.42, 80.0).round(1)
    age[rng.random(n_rows) < 0.2] = np.nan
    sibsp = rng.choice([0, 1, 2, 3, 4], size=n_rows, p=[0.68, 0.23, 0.04, 0.03, 0.02])
    parch = rng.choice([0, 1, 2, 3], size=n_rows, p=[0.76, 0.13, 0.09, 0.02])
    fare = np.round(np.exp(rng.normal(4.4 - 0.75 * pclass, 0.6, n_rows)), 4)
    embarked = rng.choice(["S", "C", "Q"], size=n_rows, p=[0.72, 0.19, 0.09]).astype(object)
    embarked[rng.choice(n_rows, size=min(2, n_rows), replace=False)] = np.nan
    cabin = np.array([f"{'ABCDEFG'[rng.integers(7)]}{rng.integers(1, 130)}" for _ in range(n_rows)], dtype=object)
    cabin[rng.random(n_rows) < 0.77] = np.nan
    logit = 1.2 - 2.5 * male - 0.8 * (pclass - 2) + 0.004 * fare - 0.01 * np.nan_to_num(age, nan=29.7)
    survived = (rng.random(n_rows) < 1.0 / (1.0 + np.exp(-logit))).astype(int)
    table = pd.DataFrame({
        "PassengerId": np.arange(1, n_rows + 1),
        "Survived": survived,
        "Pclass": pclass,
        "Name": [f"Passenger, {i}" for i in range(n_rows)],
        "Sex": np.where(male, "male", "female"),
        "Age": age,
        "SibSp": sibsp,
        "Parch": parch,
        "Ticket": [str(100000 + int(t)) for t in rng.integers(0, 900000, n_rows)],
        "Fare": fare,
        "Cabin": cabin,
        "Embarked": embarked,
    })
    table.to_csv(path, index=False)
