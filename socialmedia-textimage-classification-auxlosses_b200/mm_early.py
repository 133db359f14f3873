"""The image-text auxiliary-loss tail of the reference's EARLY-fusion models (models/mm_early.py) on the same kernels
(SURVEY.md §8 row f-3).  ViLT / LXMERT themselves are out of scope (different encoder family); what this module mirrors is
what they do with their two pooled embeddings afterwards:

    ViLT.get_logits_per_text / Lxmert.get_logits_per_text(text_embeds, image_embeds)      mm_early.py:96-103, 165-172
        L2-normalise both (no epsilon), logits_per_text = exp(logit_scale) * T̂ V̂ᵀ
    MMEarly_Model.prepare_itm_inputs(ids, mask, token_type_ids)                            mm_early.py:262-293
    the loss mix of MMEarly_Model.train                                                     mm_early.py:366-379

`get_logits_per_text` is differentiable w.r.t. BOTH embeddings and logit_scale (the early-fusion encoders are trained end to
end: there is no frozen tower here), through the tcgen05 similarity tiles and gradient GEMMs of csrc/itc.cu.  The embeddings
arrive in fp32 from the encoder; they are consumed as split bf16 (hi, lo) pairs, so nothing is lost to the bf16 rounding of
the tensor-core operands.  `itc_loss` is the fused form (clip_loss of those logits without materialising them).

No CPU path: CUDA tensors and the built library are required.
"""
import math

import numpy as np
import torch

from . import capi
from .capi import call, ptr
from .plan import ItcPlan, _stream
from .utils import clip_loss

_PLANS = {}


def _plan(B, d, dev, logits):
    key = (B, d, str(dev), bool(logits))
    if key not in _PLANS:
        if d % 8:
            raise ValueError("embedding width must be a multiple of 8")
        _PLANS[key] = ItcPlan(B, B, d, dev, materialize_logits=logits, precise=True if logits else None)
    return _PLANS[key]


def _device_scale(logit_scale, dev):
    """exp(logit_scale) as a device scalar (clamped to (0, 40] by the library, status word sticky): no host read"""
    ls = logit_scale.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
    sc = torch.empty(1, dtype=torch.float32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    call("tic_refresh_weights", 0, None, None, None, None, None, None, ptr(ls), ptr(sc), ptr(status), None, 0, None, 0, _stream())
    return sc


def _enter(it):
    """The forward activations (inverse norms, partial sums) live in the shape-keyed plan: stamp the forward so that a
    backward which follows ANOTHER forward on the same plan fails loudly instead of differentiating the wrong buffers."""
    it.generation = getattr(it, "generation", 0) + 1
    return it.generation


def _check_generation(it, gen):
    if getattr(it, "generation", 0) != gen:
        raise RuntimeError("tic_b200.mm_early: another forward with the same (B, d) ran between this forward and its backward; "
                           "the saved activations are gone (call backward before the next forward of the same shape)")


def _split(x):
    """fp32 [B, d] -> contiguous bf16 (hi, lo) with hi + lo == x to ~16 mantissa bits"""
    x = x.detach().to(torch.float32).contiguous()
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


def _check(text_embeds, image_embeds):
    if not (text_embeds.is_cuda and image_embeds.is_cuda):
        raise capi.TicError("tic_b200.mm_early needs CUDA tensors: this package has no CPU path")
    if text_embeds.dim() != 2 or text_embeds.shape != image_embeds.shape:
        raise ValueError("text_embeds and image_embeds must both be [B, d]")


class _LogitsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, text_embeds, image_embeds, logit_scale):
        _check(text_embeds, image_embeds)
        B, d = text_embeds.shape
        it = _plan(B, d, text_embeds.device, True)
        T, Tl = _split(text_embeds)
        V, Vl = _split(image_embeds)
        scale = 0.0                                  # unused: the kernels read the device scalar
        it.scale_dev = sc = _device_scale(logit_scale, T.device)
        ctx.gen = _enter(it)
        it.norms(T, d, V, d, T_lo=Tl, V_lo=Vl)
        it.fwd_tiles(T, d, V, d, scale, T_lo=Tl, V_lo=Vl)
        ctx.it, ctx.ops, ctx.scale, ctx.sc = it, (T, Tl, V, Vl), scale, sc
        ctx.dtypes = (text_embeds.dtype, image_embeds.dtype, logit_scale.dtype)
        return it.logits.clone().to(text_embeds.dtype)

    @staticmethod
    def backward(ctx, dS):
        it, (T, Tl, V, Vl), scale = ctx.it, ctx.ops, ctx.scale
        _check_generation(it, ctx.gen)
        it.scale_dev = ctx.sc
        B, d = T.shape
        dev = T.device
        dS = dS.detach().to(torch.float32).contiguous()
        dT = torch.empty(B, d, dtype=torch.float32, device=dev)
        dV = torch.empty(B, d, dtype=torch.float32, device=dev)
        r_sum = torch.zeros(1, dtype=torch.float32, device=dev)
        call("tic_itc_ds_operands", ptr(dS), dS.stride(0), B, B, ptr(it.rinv_t), ptr(it.rinv_v), ptr(it.GA), ptr(it.GA_lo),
             it.ld_ga, ptr(it.GBT), ptr(it.GBT_lo), it.ld_gbt, _stream())
        it.grad_gemms(T, d, V, d, T_lo=Tl, V_lo=Vl)
        it.finalize_t(T, d, V, d, it.rinv_v, scale, 0.0, dT, None, r_sum, T_lo=Tl, V_diag_lo=Vl)
        it.finalize_v(it.acc_v, V, d, it.rinv_v, T, d, it.rinv_t, B, scale, 0.0, dV, None, V_lo=Vl, T_diag_lo=Tl)
        dt_t, dt_v, dt_s = ctx.dtypes
        return dT.to(dt_t), dV.to(dt_v), r_sum.reshape(()).to(dt_s)


def get_logits_per_text(text_embeds, image_embeds, logit_scale):
    """mm_early.py:96-103 / :165-172 — exp(logit_scale) * normalise(text_embeds) @ normalise(image_embeds).T  ([B, B])."""
    return _LogitsFn.apply(text_embeds, image_embeds, logit_scale)


class _ItcLossFn(torch.autograd.Function):
    """clip_loss(get_logits_per_text(T, V, ls)) with the bidirectional softmax-CE fused in the tile epilogue: the [B, B]
    logits never reach HBM (forward) and the backward recomputes the tiles."""

    @staticmethod
    def forward(ctx, text_embeds, image_embeds, logit_scale):
        _check(text_embeds, image_embeds)
        B, d = text_embeds.shape
        it = _plan(B, d, text_embeds.device, False)
        lo = it.precise
        T, Tl = _split(text_embeds)
        V, Vl = _split(image_embeds)
        if not lo:            # >= 4096 negatives: the embeddings are consumed as single bf16 — the residual K-segments are
            Tl = Vl = None    # skipped where tensor time matters; below that, (hi, lo) pairs keep fp32 inputs intact
        scale = 0.0                                  # unused: the kernels read the device scalar
        it.scale_dev = sc = _device_scale(logit_scale, T.device)
        ctx.gen = _enter(it)
        sums = torch.zeros(2, dtype=torch.float32, device=T.device)
        it.norms(T, d, V, d, T_lo=Tl, V_lo=Vl)
        it.fwd_tiles(T, d, V, d, scale, T_lo=Tl, V_lo=Vl)
        it.lse_loss(scale, sums)
        ctx.it, ctx.ops, ctx.scale, ctx.sc = it, (T, Tl, V, Vl), scale, sc
        ctx.dtypes = (text_embeds.dtype, image_embeds.dtype, logit_scale.dtype)
        return (0.5 * sums.sum() / B).to(text_embeds.dtype)

    @staticmethod
    def backward(ctx, g):
        it, (T, Tl, V, Vl), scale = ctx.it, ctx.ops, ctx.scale
        _check_generation(it, ctx.gen)
        it.scale_dev = ctx.sc
        B, d = T.shape
        dev = T.device
        gf = float(g)          # upstream scalar (one host read; the fused HeadPlan path has none)
        dT = torch.empty(B, d, dtype=torch.float32, device=dev)
        dV = torch.empty(B, d, dtype=torch.float32, device=dev)
        r_sum = torch.zeros(1, dtype=torch.float32, device=dev)
        it.bwd_operands(T, d, V, d, scale, gf / (2.0 * B), T_lo=Tl, V_lo=Vl)
        it.grad_gemms(T, d, V, d, T_lo=Tl, V_lo=Vl)
        it.finalize_t(T, d, V, d, it.rinv_v, scale, gf / B, dT, None, r_sum, T_lo=Tl, V_diag_lo=Vl)
        it.finalize_v(it.acc_v, V, d, it.rinv_v, T, d, it.rinv_t, B, scale, gf / B, dV, None, V_lo=Vl, T_diag_lo=Tl)
        dt_t, dt_v, dt_s = ctx.dtypes
        return dT.to(dt_t), dV.to(dt_v), r_sum.reshape(()).to(dt_s)


def itc_loss(text_embeds, image_embeds, logit_scale):
    """utils.clip_loss(get_logits_per_text(...)) (mm_early.py:367-368 / utils.py:228-231), fused: nothing of size B² in HBM."""
    return _ItcLossFn.apply(text_embeds, image_embeds, logit_scale)


class AuxLossTail:
    """Mixin for an early-fusion model that owns `self.logit_scale` (mm_early.py:54, :126): the reference's method name."""

    def get_logits_per_text(self, text_embeds, image_embeds):
        return get_logits_per_text(text_embeds, image_embeds, self.logit_scale)


def prepare_itm_inputs(ids, mask, token_type_ids, rng="numpy", generator=None):
    """MMEarly_Model.prepare_itm_inputs (mm_early.py:262-293): as the late-fusion twin plus `token_type_ids`.  rng="numpy"
    replays the reference's global numpy stream (coin, then pick, per row); rng="device" draws the uniforms on the GPU.
    Returns fresh (tim_ids, tim_mask, tim_token_type_ids, lbl_tim)."""
    if not ids.is_cuda:
        raise capi.TicError("prepare_itm_inputs needs CUDA tensors: this package has no CPU path")
    B, dev, st = ids.shape[0], ids.device, torch.cuda.current_stream().cuda_stream
    lbl = torch.empty(B, dtype=torch.int64, device=dev)
    src = torch.empty(B, dtype=torch.int32, device=dev)
    if rng == "device":
        u = torch.rand(2, B, device=dev, generator=generator)
        call("tic_itm_sample", u[0].data_ptr(), u[1].data_ptr(), B, 0, None, 0, 0.0, None, lbl.data_ptr(), src.data_ptr(), st)
    elif rng == "numpy":
        from .mm_late import _decisions_from_numpy_stream
        swap, s_np = _decisions_from_numpy_stream(B)
        src.copy_(torch.from_numpy(s_np.astype(np.int32)))
        lbl.copy_(torch.from_numpy((~swap).astype(np.int64)))
    else:
        raise ValueError("rng must be 'numpy' or 'device'")
    outs = []
    for t in (ids, mask, token_type_ids):
        t = t.contiguous()
        o = torch.empty_like(t)                      # never aliases its inputs (mm_early.py:265-267)
        call("tic_gather_rows", t.data_ptr(), t.stride(0) * t.element_size(), o.data_ptr(), o.stride(0) * o.element_size(),
             t.shape[1] * t.element_size(), src.data_ptr(), B, st)
        outs.append(o)
    return outs[0], outs[1], outs[2], lbl


def aux_loss_mix(loss_cls, logits_per_text=None, loss_itm=None, beta_itc=0.1, beta_itm=0.1):
    """mm_early.py:366-379: (1-Σβ)·L_cls + β_itc·clip_loss(logits) + β_itm·L_itm for whichever auxiliary terms are given."""
    bi = beta_itc if logits_per_text is not None else 0.0
    bm = beta_itm if loss_itm is not None else 0.0
    loss = (1 - (bi + bm)) * loss_cls
    if logits_per_text is not None:
        loss = loss + bi * clip_loss(logits_per_text)
    if loss_itm is not None:
        loss = loss + bm * loss_itm
    return loss
