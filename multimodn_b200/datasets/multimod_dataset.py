"""Dataset contract of the fusion step (reference: datasets/multimod_dataset.py).

A sample is ``(data, targets[, encoding_sequence])`` with ``data`` a list of per-modality
tensors; default collation turns a batch into the tuple ``MultiModN.train_epoch`` consumes
(multimodn.py:119).  ``PartitionDataset.tensors()`` additionally exposes whole-partition
tensors so an epoch can live on the device instead of being rebuilt row by row.
"""
from abc import ABC, abstractmethod
from itertools import accumulate
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import Generator, Tensor
from torch.utils.data import Dataset, Subset


class MultiModDataset(Dataset, ABC):
    @abstractmethod
    def __len__(self) -> int:
        ...

    def random_split(self, probabilities: Union[List[float], Tuple[float, ...]], seed: int,
                     balanced_target_idx: Optional[int] = None) -> List[Subset]:
        """Seeded split, optionally stratified on one target (multimod_dataset.py:15-52): same
        permutation, same per-class quota rule (remainder goes to the first split)."""
        order = torch.randperm(len(self), generator=Generator().manual_seed(seed)).tolist()
        if balanced_target_idx is None:
            strata = {"all": order}
        else:
            strata = {}
            for idx in order:
                strata.setdefault(self[idx][1][balanced_target_idx], []).append(idx)
        total = sum(probabilities)
        parts: List[List[int]] = [[] for _ in probabilities]
        for members in strata.values():
            quota = [int(len(members) * p / total) for p in probabilities]
            quota[0] += len(members) - sum(quota)
            start = 0
            for part, n in zip(parts, quota):
                part.extend(members[start:start + n])
                start += n
        return [Subset(self, part) for part in parts]


class PartitionDataset(MultiModDataset):
    """Tabular matrix split column-wise into modalities (multimod_dataset.py:55-88)."""

    def __init__(self, X: np.ndarray, y: np.ndarray, partitions: Optional[Sequence[int]] = None):
        self.partitions = [X.shape[1]] if partitions is None else list(partitions)
        if sum(self.partitions) != X.shape[1]:
            raise ValueError("Paritions sum doesn't match data dimension. Expected: {}, got: {}"
                             .format(sum(self.partitions), X.shape[1]))
        self.n_partitions = len(self.partitions)
        self.X = np.split(X, list(accumulate(self.partitions[:-1])), axis=1)
        self.y = y

    def __len__(self) -> int:
        return len(self.y)

    def __getitem__(self, idx: int) -> Tuple[List[Tensor], np.ndarray]:
        return [torch.as_tensor(part[idx], dtype=torch.float32) for part in self.X], self.y[idx]

    def tensors(self, indices: Optional[Sequence[int]] = None) -> Tuple[List[Tensor], Tensor]:
        """Whole partitions as tensors (one (N, F_i) fp32 tensor per modality, targets (N, D))."""
        sel = slice(None) if indices is None else np.asarray(indices)
        data = [torch.as_tensor(np.ascontiguousarray(part[sel]), dtype=torch.float32) for part in self.X]
        return data, torch.as_tensor(np.asarray(self.y)[sel])


class FeatureWiseDataset(PartitionDataset):
    """Every column its own modality (multimod_dataset.py:91-95)."""

    def __init__(self, X: np.ndarray, y: np.ndarray):
        super().__init__(X, y, [1] * X.shape[1])


class JointDatasets(MultiModDataset):
    """Row-aligned datasets, each collapsed to one modality (multimod_dataset.py:98-114)."""

    def __init__(self, datasets: List[Dataset]):
        assert all(len(ds) == len(datasets[0]) for ds in datasets), "Datasets must have the same length"
        self.datasets = datasets

    def __len__(self) -> int:
        return len(self.datasets[0])

    def __getitem__(self, idx: int) -> Tuple[List[Tensor], np.ndarray]:
        return [torch.cat(ds[idx][0]) for ds in self.datasets], self.datasets[0][idx][1]
