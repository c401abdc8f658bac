"""End-of-test metric suite (multimodn_b200/metrics.py; reference multimodn/multimodn.py:22-49 via torchmetrics, which is
absent here): the built-in sort-based implementation is pinned against scikit-learn on the CPU, and on the GPU the device
computation must equal the CPU one."""
import numpy as np
import pytest
import torch
from sklearn.metrics import confusion_matrix, f1_score, roc_auc_score, roc_curve, precision_recall_curve

from multimodn_b200.metrics import _builtin, get_performance_metrics, performance_metrics


def _case(n=5000, seed=0, ties=True):
    g = torch.Generator().manual_seed(seed)
    y = (torch.rand(n, generator=g) < 0.35).long()
    prob = torch.sigmoid(torch.randn(n, generator=g) + 1.2 * y.float() - 0.4)
    if ties:
        prob = torch.round(prob * 200) / 200          # many equal scores: the curves must group them like sklearn does
    pred = (prob > 0.5).long()
    return y, pred, prob


@pytest.mark.parametrize("ties", [False, True])
def test_builtin_matches_sklearn(ties):
    y, pred, prob = _case(ties=ties)
    out = dict(zip(performance_metrics, _builtin(y, pred, prob)))
    yn, pn, sn = y.numpy(), pred.numpy(), prob.numpy().astype(np.float64)
    assert abs(float(out["auc"]) - roc_auc_score(yn, sn)) < 1e-6
    assert abs(float(out["f1"]) - f1_score(yn, (sn > 0.5).astype(int))) < 1e-6
    tn, fp, fn, tp = confusion_matrix(yn, pn).ravel()
    assert (int(out["tn"]), int(out["fp"]), int(out["fn"]), int(out["tp"])) == (tn, fp, fn, tp)
    assert abs(float(out["sensitivity"]) - tp / (tp + fn)) < 1e-7 and abs(float(out["specificity"]) - tn / (tn + fp)) < 1e-7
    fpr, tpr, _ = roc_curve(yn, sn, drop_intermediate=False)
    np.testing.assert_allclose(out["fpr"].numpy(), fpr, atol=1e-12)
    np.testing.assert_allclose(out["tpr"].numpy(), tpr, atol=1e-12)
    prec, rec, thr = precision_recall_curve(yn, sn)
    # sklearn drops the points after full recall is first reached; ours keeps one point per distinct score
    k = len(prec) - 1
    np.testing.assert_allclose(out["precision"].numpy()[-(k + 1):], prec, atol=1e-12)
    np.testing.assert_allclose(out["recall"].numpy()[-(k + 1):], rec, atol=1e-12)
    np.testing.assert_allclose(out["thr_pr"].numpy()[-k:], thr, atol=1e-7)


def test_non_binary_targets_give_nan():
    y = torch.tensor([0, 1, 2, 1])
    res = get_performance_metrics(y, y, torch.rand(4))
    assert all(torch.isnan(torch.as_tensor(v)).all() for v in res)


@pytest.mark.gpu
def test_device_metrics_equal_host_metrics():
    y, pred, prob = _case(n=200_000, seed=3)
    host = _builtin(y, pred, prob)
    dev = _builtin(y.cuda(), pred.cuda(), prob.cuda())
    for name, a, b in zip(performance_metrics, host, dev):
        b = b.cpu() if torch.is_tensor(b) else b
        assert torch.is_tensor(b) or isinstance(b, (int, float))
        np.testing.assert_allclose(np.asarray(torch.as_tensor(a), dtype=np.float64), np.asarray(torch.as_tensor(b), dtype=np.float64),
                                   rtol=1e-6, atol=1e-9, err_msg=name)
