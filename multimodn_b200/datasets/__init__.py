from .multimod_dataset import MultiModDataset, PartitionDataset, FeatureWiseDataset, JointDatasets

__all__ = ["MultiModDataset", "PartitionDataset", "FeatureWiseDataset", "JointDatasets"]
