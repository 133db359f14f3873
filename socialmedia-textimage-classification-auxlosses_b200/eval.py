"""Evaluation bookkeeping on the device.

The reference's eval loop (models/mm_late.py:594-612) synchronises with the host once per batch (`loss.item()`,
`(pred == target).cpu().numpy()`), grows Python lists of 0-d tensors and feeds them to six torchmetrics objects afterwards
(models/utils.py:294-325).  `EvalAccumulator` keeps predictions, targets, a confusion matrix, the loss sum and the per-batch
accuracy sum in device memory — one launch per batch (`tic_eval_accumulate`), no host synchronisation until `result()` —
and `compute_metrics` reduces the confusion matrix on the device (`tic_metrics_from_confusion`).

No CPU path: CUDA tensors and the built library are required.
"""
from typing import Dict, Optional

import torch

from . import capi
from .capi import call, ptr

METRIC_NAMES = ["f1_weighted", "f1_macro", "precision_weighted", "precision_macro", "recall_weighted", "recall_macro"]


def _stream():
    return torch.cuda.current_stream().cuda_stream


class EvalAccumulator:
    """Device-side state of one evaluation epoch over at most `capacity` samples with `num_classes` classes."""

    def __init__(self, num_classes: int, capacity: int, device="cuda"):
        self.C, self.capacity, self.dev = int(num_classes), int(capacity), torch.device(device)
        words = capi.load().tic_eval_state_words(self.C)
        if words <= 0:
            raise ValueError("num_classes must be positive")
        self.state = torch.zeros(words, dtype=torch.int64, device=self.dev)     # [C*C confusion | correct | rows | batches | scratch]
        self.sums = torch.zeros(2, dtype=torch.float32, device=self.dev)         # [sum of batch losses, sum of batch accuracies %]
        self.preds = torch.empty(self.capacity, dtype=torch.int64, device=self.dev)
        self.targets = torch.empty(self.capacity, dtype=torch.int64, device=self.dev)
        self.data_ids = None
        self.n = 0

    def reset(self):
        self.state.zero_()
        self.sums.zero_()
        self.n = 0

    def update(self, output: torch.Tensor, label: torch.Tensor, loss: Optional[torch.Tensor] = None,
               data_id: Optional[torch.Tensor] = None):
        """One batch: `output` [B, C] logits, `label` float one-hot [B, C] (mm_late.py:601) or int64 class ids [B],
        `loss` optional 0-d device tensor (the batch loss, mm_late.py:594)."""
        if not output.is_cuda:
            raise capi.TicError("EvalAccumulator needs CUDA tensors: this package has no CPU path")
        B, C = output.shape
        if C != self.C:
            raise ValueError("expected %d classes, got %d" % (self.C, C))
        if self.n + B > self.capacity:
            raise ValueError("EvalAccumulator capacity %d exceeded" % self.capacity)
        out = output.detach().to(torch.float32).contiguous()
        y_soft = y_int = None
        if label.dtype in (torch.int64, torch.int32) and label.dim() == 1:
            y_int = label.to(torch.int64).contiguous()
        else:
            y_soft = label.detach().to(torch.float32).contiguous()
        lossf = None if loss is None else loss.detach().to(torch.float32).reshape(1).contiguous()
        call("tic_eval_accumulate", ptr(out), out.stride(0), ptr(y_soft), 0 if y_soft is None else y_soft.stride(0),
             ptr(y_int), B, C, ptr(lossf), self.preds.data_ptr() + 8 * self.n, self.targets.data_ptr() + 8 * self.n,
             ptr(self.state), ptr(self.sums), _stream())
        if data_id is not None:
            if self.data_ids is None:
                self.data_ids = torch.empty((self.capacity,) + tuple(data_id.shape[1:]), dtype=data_id.dtype, device=self.dev)
            self.data_ids[self.n:self.n + B].copy_(data_id)
        self.n += B

    @property
    def confusion(self) -> torch.Tensor:
        """[C, C] int64, row = target, column = prediction (a view of the device state)."""
        return self.state[: self.C * self.C].view(self.C, self.C)

    def metrics_device(self) -> torch.Tensor:
        """fp32 [6] on the device, in METRIC_NAMES order (utils.py:294-325, single-label branch)."""
        out = torch.empty(6, dtype=torch.float32, device=self.dev)
        call("tic_metrics_from_confusion", ptr(self.state), self.C, ptr(out), _stream())
        return out

    def result(self) -> Dict:
        """The reference's eval() return value (mm_late.py:629-636) — ONE host synchronisation for the whole epoch."""
        counts = self.state[self.C * self.C:].tolist()       # [correct, rows, batches, scratch]  (synchronises)
        nb = max(int(counts[2]), 1)
        sums = self.sums.tolist()
        res = {"loss": sums[0] / nb, "accuracy": sums[1] / nb, "predictions": self.preds[: self.n], "labels": self.targets[: self.n]}
        if self.data_ids is not None:
            res["data_id"] = self.data_ids[: self.n]
        return res


def compute_metrics(res: Dict, num_classes: int, multi_label: bool = False) -> Dict[str, list]:
    """models/utils.py:294-325: {"metric": [...six names..., "loss"], "result": [...]} from an eval() result dict.  The
    confusion matrix and the six scores are computed on the device; the values leave it in one copy."""
    if multi_label:
        raise NotImplementedError("multilabel metrics belong to the unreachable task 10 of the reference (config.py:10)")
    pred, tgt = res["predictions"], res["labels"]
    if not pred.is_cuda:
        raise capi.TicError("compute_metrics needs CUDA tensors: this package has no CPU path")
    acc = EvalAccumulator(num_classes, max(int(pred.numel()), 1), device=pred.device)
    # predictions are already class ids: feed them as one-hot "logits" of a single pass over the whole epoch
    onehot = torch.nn.functional.one_hot(pred.to(torch.int64).reshape(-1), num_classes).to(torch.float32)
    acc.update(onehot, tgt.to(torch.int64).reshape(-1))
    vals = acc.metrics_device().tolist()
    return {"metric": METRIC_NAMES + ["loss"], "result": vals + [res["loss"]]}


def agg_metrics_val(res_val, metric_names, num_labels):
    """models/utils.py:327-336."""
    out = {"metric": metric_names}
    for predictions in res_val:
        m = compute_metrics(predictions, num_labels)
        d = dict(zip(m["metric"], m["result"]))
        out["epoch-" + str(predictions["epoch"] + 1)] = [d[k] for k in metric_names]
    return out
