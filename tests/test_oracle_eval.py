"""CPU: the evaluation-bookkeeping restatement (oracle/restatement.py: eval_batch, confusion_matrix, metrics_from_confusion,
compute_metrics — models/mm_late.py:596-608, models/utils.py:294-325).  torchmetrics (pinned 0.11.0 by the reference) is not
installed, so the torchmetrics semantics are pinned against scikit-learn where the two coincide and against hand-computed
known answers where they differ (classes that never occur)."""
import numpy as np
import pytest
import torch

from oracle import restatement as R


@pytest.mark.parametrize("C,n,seed", [(2, 50, 0), (3, 301, 1), (4, 1000, 2), (7, 5000, 3)])
def test_metrics_match_sklearn_when_every_class_occurs(C, n, seed):
    from sklearn.metrics import f1_score, precision_score, recall_score
    rs = np.random.RandomState(seed)
    t = rs.randint(0, C, n)
    p = np.where(rs.rand(n) < 0.55, t, rs.randint(0, C, n))
    assert len(set(t)) == C and len(set(p)) == C
    m = R.metrics_from_confusion(R.confusion_matrix(p, t, C))
    for avg in ("weighted", "macro"):
        assert abs(m["f1_" + avg] - f1_score(t, p, average=avg)) < 1e-12
        assert abs(m["precision_" + avg] - precision_score(t, p, average=avg)) < 1e-12
        assert abs(m["recall_" + avg] - recall_score(t, p, average=avg)) < 1e-12


def test_metrics_known_answers():
    # perfect prediction
    m = R.metrics_from_confusion(np.diag([5, 3, 2]))
    assert all(abs(v - 1.0) < 1e-12 for v in m.values())
    # class 2 never occurs as target nor prediction: dropped from `macro` (torchmetrics 0.11), weight 0 in `weighted`
    conf = np.array([[3, 1, 0], [2, 4, 0], [0, 0, 0]])
    m = R.metrics_from_confusion(conf)
    p0, p1 = 3 / 5, 4 / 5
    r0, r1 = 3 / 4, 4 / 6
    assert abs(m["precision_macro"] - (p0 + p1) / 2) < 1e-12 and abs(m["recall_macro"] - (r0 + r1) / 2) < 1e-12
    assert abs(m["recall_weighted"] - (4 * r0 + 6 * r1) / 10) < 1e-12
    # class 1 is predicted but never a target: precision 0, recall 0/0 -> 0, still counted by `macro` (tp+fp+fn > 0)
    conf = np.array([[2, 2], [0, 0]])
    m = R.metrics_from_confusion(conf)
    assert abs(m["precision_macro"] - (1.0 + 0.0) / 2) < 1e-12 and abs(m["recall_macro"] - (0.5 + 0.0) / 2) < 1e-12
    assert abs(m["f1_weighted"] - (2 * 2 / (2 * 2 + 0 + 2))) < 1e-12
    # nothing at all
    assert all(v == 0.0 for v in R.metrics_from_confusion(np.zeros((3, 3))).values())


def test_eval_batch_follows_the_reference_lines():
    out = torch.tensor([[0.1, 2.0, -1.0], [3.0, 3.0, 0.0], [-5.0, -6.0, -4.0]])
    lab = torch.tensor([[0.0, 1.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    pred, tgt, acc = R.eval_batch(out, lab)
    assert pred.tolist() == [1, 0, 2]          # exact tie -> first index (torch.argmax)
    assert tgt.tolist() == [1, 1, 2] and abs(acc - 200.0 / 3) < 1e-9
    res = {"predictions": pred, "labels": tgt, "loss": 0.5}
    cm = R.compute_metrics(res, 3)
    assert cm["metric"][-1] == "loss" and cm["result"][-1] == 0.5 and len(cm["metric"]) == 7
