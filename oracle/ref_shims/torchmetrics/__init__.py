"""Stub of torchmetrics for importing the reference (see ../README.md)."""
import torch


class ConfusionMatrix:
    def __init__(self, task="binary", num_classes=2, **kw):
        if task != "binary":
            raise NotImplementedError("stub: binary only")

    def to(self, device):
        return self

    def __call__(self, preds, target):
        preds = preds.long().flatten()
        target = target.long().flatten()
        return torch.bincount(target * 2 + preds, minlength=4).reshape(2, 2)


class _NaNMetric:
    def __init__(self, *a, **kw):
        pass

    def to(self, device):
        return self

    def __call__(self, *a, **kw):
        return torch.tensor(float("nan"))


class F1Score(_NaNMetric):
    pass


class Accuracy(_NaNMetric):
    pass


class AUROC(_NaNMetric):
    pass


class _NaNCurve(_NaNMetric):
    def __call__(self, *a, **kw):
        nan = torch.tensor([float("nan")])
        return nan, nan, nan


class ROC(_NaNCurve):
    pass


class PrecisionRecallCurve(_NaNCurve):
    pass
