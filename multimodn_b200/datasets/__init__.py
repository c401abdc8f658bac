from .multimod_dataset import MultiModDataset, PartitionDataset, FeatureWiseDataset, JointDatasets
from .titanic import TitanicDataset, write_synthetic_titanic_csv

__all__ = ["MultiModDataset", "PartitionDataset", "FeatureWiseDataset", "JointDatasets", "TitanicDataset",
           "write_synthetic_titanic_csv"]
