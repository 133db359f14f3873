"""Mirror of the functions of the reference's models/utils.py that sit on the hot path: `contrastive_loss`, `clip_loss`
(utils.py:225-231) and `get_optimizer_params` (utils.py:280-292); `compute_metrics` / `agg_metrics_val` (utils.py:294-336)
are the device-side versions of tic_b200.eval, imported lazily below.  `clip_loss` runs the hand-written bidirectional
softmax-CE kernels (csrc/ce.cu) through the C ABI and is differentiable (torch.autograd.Function)."""
import torch

from . import capi
from .capi import call


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _ClipLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, similarity):
        if not similarity.is_cuda:
            raise capi.TicError("tic_b200.utils.clip_loss needs a CUDA tensor: this package has no CPU path")
        if similarity.dim() != 2 or similarity.shape[0] != similarity.shape[1]:
            raise ValueError("clip_loss expects a square [B,B] similarity matrix (in-batch positives on the diagonal)")
        S = similarity.detach().to(torch.float32).contiguous()
        B = S.shape[0]
        dev = S.device
        lse_row = torch.empty(B, dtype=torch.float32, device=dev)
        lse_col = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        ws = torch.empty(capi.load().tic_ce_bidir_workspace_bytes(B), dtype=torch.uint8, device=dev)
        call("tic_ce_bidir_fwd", S.data_ptr(), S.stride(0), B, lse_row.data_ptr(), lse_col.data_ptr(), loss.data_ptr(),
             ws.data_ptr(), _stream())
        ctx.save_for_backward(S, lse_row, lse_col)
        ctx.in_dtype = similarity.dtype
        return loss[0].to(similarity.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        S, lse_row, lse_col = ctx.saved_tensors
        B = S.shape[0]
        g = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        dS = torch.empty_like(S)
        call("tic_ce_bidir_bwd", S.data_ptr(), S.stride(0), B, lse_row.data_ptr(), lse_col.data_ptr(), g.data_ptr(),
             dS.data_ptr(), dS.stride(0), _stream())
        return dS.to(ctx.in_dtype)


def clip_loss(similarity: torch.Tensor) -> torch.Tensor:
    """models/utils.py:228-231 — (CE(S, arange) + CE(S^T, arange)) / 2, one read of S forward, one read + one write backward."""
    return _ClipLossFn.apply(similarity)


def contrastive_loss(logits: torch.Tensor) -> torch.Tensor:
    """models/utils.py:225-226 — F.cross_entropy(logits, arange(B)).  Kept as a library call: off the hot path once
    clip_loss is fused (the reference only ever calls it through clip_loss)."""
    return torch.nn.functional.cross_entropy(logits, torch.arange(len(logits), device=logits.device))


def get_optimizer_params(named_parameters, weight_decay, lr, verbose=False):
    """models/utils.py:280-292 — one param group with every requires_grad parameter."""
    params = {"lr": lr, "weight_decay": weight_decay, "params": []}
    for name, param in named_parameters:
        if verbose:
            print(name)
        if param.requires_grad:
            params["params"].append(param)
    return [params]


def compute_metrics(res, num_classes, multi_label=False):
    """models/utils.py:294-325 (see tic_b200.eval.compute_metrics: confusion matrix + six scores on the device)."""
    from .eval import compute_metrics as _cm
    return _cm(res, num_classes, multi_label=multi_label)


def agg_metrics_val(res_val, metric_names, num_labels):
    """models/utils.py:327-336."""
    from .eval import agg_metrics_val as _agg
    return _agg(res_val, metric_names, num_labels)
