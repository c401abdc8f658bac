"""End-of-test metric suite on the last-encoder outputs (reference: multimodn/multimodn.py:18-49).

Post-processing of ``(N,)`` probabilities, outside the fused step.  It runs on
the device the probabilities live on: ``MultiModN.test`` keeps the collected last-step outputs and targets in HBM, the
sort-based ROC / PR curves, AUROC, F1 and the confusion cells are torch CUDA ops there, and only the resulting tuple
crosses PCIe.  The reference delegates to torchmetrics, which is unpinned there and absent here; when it is importable
it is used, otherwise the same quantities are computed below (pinned against scikit-learn by
tests/test_metrics.py).  The tuple layout is the reference's ``performance_metrics`` list.
"""
import torch

performance_metrics = ['f1', 'auc', 'accuracy', 'sensitivity', 'specificity', 'fpr', 'tpr', 'precision', 'recall',
                       'tn', 'fp', 'fn', 'tp', 'thr_roc', 'thr_pr']


def _curves(y_true, y_prob):
    """ROC and precision-recall points at every distinct score, scores descending."""
    order = torch.argsort(y_prob, descending=True, stable=True)
    score, truth = y_prob[order], y_true[order].to(torch.float64)
    distinct = torch.nonzero(score[1:] != score[:-1]).flatten()
    idx = torch.cat([distinct, torch.tensor([score.numel() - 1], device=score.device)])
    tps = torch.cumsum(truth, 0)[idx]
    fps = (idx + 1).to(torch.float64) - tps
    thr = score[idx]
    pos, neg = truth.sum(), truth.numel() - truth.sum()
    zero = torch.zeros(1, dtype=torch.float64, device=score.device)
    tpr = torch.cat([zero, tps / pos if pos > 0 else torch.zeros_like(tps)])
    fpr = torch.cat([zero, fps / neg if neg > 0 else torch.zeros_like(fps)])
    thr_roc = torch.cat([torch.ones(1, dtype=thr.dtype, device=thr.device), thr])
    precision = torch.flip(torch.cat([zero + 1, tps / (tps + fps)])[1:], [0])
    recall = torch.flip((tps / pos if pos > 0 else torch.zeros_like(tps)), [0])
    precision = torch.cat([precision, zero + 1])
    recall = torch.cat([recall, zero])
    return fpr, tpr, thr_roc, precision, recall, torch.flip(thr, [0])


def _builtin(y_true, y_pred, y_prob):
    y_true = y_true.long().flatten()
    y_pred = y_pred.long().flatten()
    y_prob = y_prob.flatten().to(torch.float32)
    cm = torch.bincount(y_true * 2 + y_pred, minlength=4).reshape(2, 2)
    tn, fp, fn, tp = cm[0][0], cm[0][1], cm[1][0], cm[1][1]
    sensitivity = tp / (tp + fn) if (tp + fn) != 0 else 0
    specificity = tn / (tn + fp) if (tn + fp) != 0 else 0
    hard = (y_prob > 0.5).long()
    tp_h = ((hard == 1) & (y_true == 1)).sum()
    fp_h = ((hard == 1) & (y_true == 0)).sum()
    fn_h = ((hard == 0) & (y_true == 1)).sum()
    f1 = 2 * tp_h / (2 * tp_h + fp_h + fn_h) if (2 * tp_h + fp_h + fn_h) != 0 else torch.tensor(0.0, device=y_prob.device)
    fpr, tpr, thr_roc, precision, recall, thr_pr = _curves(y_true, y_prob)
    auc = torch.trapz(tpr, fpr).to(torch.float32)
    accuracy = (y_pred == y_true).float().mean()
    return (f1, auc, accuracy, sensitivity, specificity, fpr, tpr, precision, recall, tn, fp, fn, tp, thr_roc, thr_pr)


def to_host(result):
    """the metric tuple with every tensor moved to the host (what the reference's callers index and print)"""
    return tuple(v.cpu() if torch.is_tensor(v) else v for v in result)


def get_performance_metrics(y_true, y_pred, y_prob):
    if int(y_true.max()) > 1 or int(y_pred.max()) > 1:
        nan = torch.tensor(float("nan"))
        return tuple(nan for _ in performance_metrics)        # the reference supports binary tasks only
    try:
        from torchmetrics import ConfusionMatrix, F1Score, ROC, PrecisionRecallCurve, Accuracy, AUROC
    except ImportError:
        return _builtin(y_true, y_pred, y_prob)
    cm = ConfusionMatrix(task="binary")(y_pred, y_true)
    tp, fp, fn, tn = cm[1][1], cm[0][1], cm[1][0], cm[0][0]
    sensitivity = tp / (tp + fn) if (tp + fn) != 0 else 0
    specificity = tn / (tn + fp) if (tn + fp) != 0 else 0
    fpr, tpr, thr_roc = ROC(task="binary")(y_prob, y_true)
    precision, recall, thr_pr = PrecisionRecallCurve(task="binary")(y_prob, y_true)
    return (F1Score(task="binary", average='macro')(y_prob, y_true), AUROC(task="binary", average='macro')(y_prob, y_true),
            Accuracy(task="binary")(y_pred, y_true), sensitivity, specificity, fpr, tpr, precision, recall,
            tn, fp, fn, tp, thr_roc, thr_pr)
