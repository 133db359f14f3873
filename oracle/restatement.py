"""TEST INFRASTRUCTURE — CPU restatement (the parity oracle) of the reference's late-fusion head and image-text
auxiliary-loss path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module; the product package never does (it fails loudly when the CUDA library is missing).

Every function cites the reference lines it restates (paths relative to /root/reference; `HF:` is
transformers/models/vision_text_dual_encoder/modeling_vision_text_dual_encoder.py, pinned ==4.25.1 by
timrel-env.yml:122, a third-party dependency that holds the ITC logits arithmetic for mm_late).

Pinning: the reference ships no tests / golden vectors for this path.  This restatement is
pinned against the reference's own code run in the build container under the shims of oracle/ref_shims.py
(tests/test_oracle_vs_reference.py) and against fixtures generated from it by oracle/make_golden.py
(tests/golden/*.npz).  The similarity-weighted hard-negative sampler (itm_sample_hard) is an extension named only
by BASELINE.json: its parity is UNPINNED by the reference — the spec is this file.  metrics_from_confusion restates
torchmetrics==0.11.0 (third-party, absent here): PARITY UNPINNED against torchmetrics, pinned against scikit-learn.

Floating point is torch CPU (fp32 or fp64, chosen by the dtype of the inputs); integer / RNG work is numpy.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------------- ITC


def project(pool: torch.Tensor, weight: Optional[torch.Tensor]) -> torch.Tensor:
    """HF:261-265 — text_projection / visual_projection are nn.Linear(E, P, bias=False). weight [P,E]."""
    return pool if weight is None else pool @ weight.t()


def l2_normalize(x: torch.Tensor) -> torch.Tensor:
    """HF:268-269; models/mm_early.py:98-99 — x / ||x||_2 along the last dim, no epsilon."""
    return x / x.norm(p=2, dim=-1, keepdim=True)


def itc_logits(text_embeds: torch.Tensor, image_embeds: torch.Tensor, logit_scale: torch.Tensor) -> torch.Tensor:
    """HF:272-273; models/mm_early.py:96-103 — logits_per_text = exp(logit_scale) * That @ Vhat^T."""
    image_embeds = l2_normalize(image_embeds)
    text_embeds = l2_normalize(text_embeds)
    return torch.matmul(text_embeds, image_embeds.t()) * logit_scale.exp()


def contrastive_loss(logits: torch.Tensor) -> torch.Tensor:
    """models/utils.py:225-226."""
    return F.cross_entropy(logits, torch.arange(len(logits), device=logits.device))


def clip_loss(similarity: torch.Tensor) -> torch.Tensor:
    """models/utils.py:228-231."""
    return (contrastive_loss(similarity) + contrastive_loss(similarity.t())) / 2.0


# ---------------------------------------------------------------------------------------------------- fusion


def sdpa(Q, K, V, scale=None):
    """models/mm_late.py:199-210 — returns (context, raw attention scores before scaling)."""
    attention = torch.matmul(Q, K.permute(0, 2, 1))
    scores = attention
    if scale:
        attention = attention * scale
    attention = F.softmax(attention, dim=-1)
    return torch.matmul(attention, V), scores


def _relu(x, mask=None, pre_out=None):
    """self.relu (mm_late.py:82).  `mask` (test-only) replaces the comparison x > 0 by a given 0/1 mask so gradients can be
    compared with a lower-precision implementation whose near-zero pre-activations land on the other side of 0;
    `pre_out` (a list) receives the pre-activation."""
    if pre_out is not None:
        pre_out.append(x.detach())
    return F.relu(x) if mask is None else x * mask.to(x.dtype)


def fusion_concat(xt_cls, xv_cls, p, relu_mask=None, pre_out=None):
    """models/mm_late.py:92-96."""
    return _relu(F.linear(torch.cat((xt_cls, xv_cls), dim=1), p["linear_fusion.weight"], p["linear_fusion.bias"]),
                 relu_mask, pre_out)


def fusion_attention_literal(x_t, x_v, p, relu_mask=None, pre_out=None):
    """models/mm_late.py:98-113 exactly as written (all Lt query rows computed, only row 0 used)."""
    N, L, E = x_t.shape
    Q = F.linear(x_t, p["fc_Q.weight"], p["fc_Q.bias"])
    K = F.linear(x_v, p["fc_K.weight"], p["fc_K.bias"])
    V = F.linear(x_v, p["fc_V.weight"], p["fc_V.bias"])
    ctx, _ = sdpa(Q, K, V, K.size(-1) ** -0.5)
    ctx = ctx.view(N, L, E)
    return fusion_concat(x_t[:, 0, :], ctx[:, 0, :], p, relu_mask, pre_out)


def fusion_attention_collapsed(xt_cls, x_v, p, relu_mask=None, pre_out=None):
    """The exact CLS-row algebraic collapse of mm_late.py:98-113, which is what the CUDA path
    computes: only query row 0 reaches the output (mm_late.py:111)."""
    E = x_v.shape[-1]
    q0 = F.linear(xt_cls, p["fc_Q.weight"], p["fc_Q.bias"])          # [B,E]
    kq = q0 @ p["fc_K.weight"]                                        # W_K^T q0   [B,E]
    c = q0 @ p["fc_K.bias"]                                           # [B]
    s = (torch.einsum("be,ble->bl", kq, x_v) + c[:, None]) * (E ** -0.5)
    a = F.softmax(s, dim=-1)
    xbar = torch.einsum("bl,ble->be", a, x_v)
    ctx0 = F.linear(xbar, p["fc_V.weight"], p["fc_V.bias"])
    return fusion_concat(xt_cls, ctx0, p, relu_mask, pre_out)


def fusion_aspect(t_pool, v_pool, p, relu_mask=None, pre_out=None):
    """models/mm_late.py:115-131 including the stack -> reshape (not transpose) row scrambling :120-121:
    sample i pairs flat rows 2i, 2i+1 of [t_0..t_{B-1}, v_0..v_{B-1}]."""
    N, EM = t_pool.shape
    V = torch.stack((t_pool, v_pool), dim=0)
    V = torch.reshape(V, (N, 2, EM))
    Ew = torch.tanh(F.linear(V, p["aspectattention.weight"], p["aspectattention.bias"]))
    w = F.softmax(Ew, dim=1).transpose(1, 2)
    return _relu(torch.matmul(w, V).squeeze(1), relu_mask, pre_out)


def fusion_gmu(xt_cls, xv_cls, p, relu_mask=None, pre_out=None):
    """models/mm_late.py:133-144 — the gate is sigmoid of the raw concatenation (no learned gate matrix)."""
    v_prime = F.linear(xv_cls, p["linear_gmu_v.weight"], p["linear_gmu_v.bias"])
    t_prime = F.linear(xt_cls, p["linear_gmu_t.weight"], p["linear_gmu_t.bias"])
    z = torch.sigmoid(torch.cat((xt_cls, xv_cls), dim=1))
    g = z * t_prime + (1 - z) * v_prime
    return _relu(F.linear(g, p["linear_fusion.weight"], p["linear_fusion.bias"]), relu_mask, pre_out)


def mm_fusion(fusion_name, x_t, x_v, p, x_v_pool=None, x_t_pool=None, literal_attention=False, relu_mask=None,
              pre_out=None):
    """models/mm_late.py:91-144 dispatch. x_t [B,Lt,E], x_v [B,Lv,E]."""
    if fusion_name == "concat":
        return fusion_concat(x_t[:, 0, :], x_v[:, 0, :], p, relu_mask, pre_out)
    if fusion_name == "attention":
        if literal_attention:
            return fusion_attention_literal(x_t, x_v, p, relu_mask, pre_out)
        return fusion_attention_collapsed(x_t[:, 0, :], x_v, p, relu_mask, pre_out)
    if fusion_name == "aspect-att":
        if x_t_pool is None or x_v_pool is None:
            # mm_late.py:181 calls mm_fusion without pools on the ITM branch -> torch.stack((None, None)) TypeError
            raise TypeError("aspect-att needs pooled outputs (the reference crashes here when ITM is on)")
        return fusion_aspect(x_t_pool, x_v_pool, p, relu_mask, pre_out)
    if fusion_name == "gmu":
        return fusion_gmu(x_t[:, 0, :], x_v[:, 0, :], p, relu_mask, pre_out)
    return None  # mm_late.py falls through


# ---------------------------------------------------------------------------------------------------- losses


def cls_loss_soft(logits, y_soft, class_w=None):
    """run_mm_late.py:85 + mm_late.py:471,474: CrossEntropyLoss(weight=w)(logits, float one-hot labels) =
    -(1/B) sum_i sum_c w_c y_ic log softmax(logits)_ic  (probability-target form: divides by B, not sum w)."""
    return F.cross_entropy(logits, y_soft, weight=class_w)


def itm_loss(logits, labels):
    """run_mm_late.py:97: CrossEntropyLoss()(out_tim, int64 labels)."""
    return F.cross_entropy(logits, labels)


def loss_mix(l_cls, l_itc, l_itm, use_itc, use_itm, beta_itc, beta_itm):
    """models/mm_late.py:473-487 (train) / :581-593 (eval)."""
    if use_itc and use_itm:
        return (1 - (beta_itc + beta_itm)) * l_cls + beta_itc * l_itc + beta_itm * l_itm
    if use_itc:
        return (1 - beta_itc) * l_cls + beta_itc * l_itc
    if use_itm:
        return (1 - beta_itm) * l_cls + beta_itm * l_itm
    return l_cls


# ---------------------------------------------------------------------------------------------------- ITM sampling


def itm_decisions_from_stream(B: int, rng) -> Tuple[np.ndarray, np.ndarray]:
    """models/mm_late.py:395-409 on the legacy numpy stream `rng` (np.random module or a RandomState):
    returns swap[B] (bool) and src[B] (row whose text ends up at position i)."""
    swap = np.zeros(B, dtype=bool)
    src = np.arange(B, dtype=np.int64)
    if B > 1:
        for idx in range(B):
            if rng.choice([True, False]):
                swap[idx] = True
                indexes = set(range(B)) - {idx}
                src[idx] = rng.choice(list(indexes))
    return swap, src


def prepare_itm_inputs_stream(ids: torch.Tensor, mask: torch.Tensor, rng):
    """models/mm_late.py:389-414 — clones, per-row coin + uniform other-row pick, labels 0 = mismatch, 1 = match."""
    swap, src = itm_decisions_from_stream(ids.shape[0], rng)
    idx = torch.from_numpy(src)
    return ids[idx].clone(), mask[idx].clone(), torch.from_numpy((~swap).astype(np.int64))


def itm_sample_uniform(u_coin: np.ndarray, u_pick: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Uniform-driven statement of mm_late.py:395-409 (what tic_itm_sample computes in TIC_ITM_UNIFORM mode):
    swap iff u_coin < 0.5; k = min(floor(fp32(u_pick * (B-1))), B-2); src = k if k < i else k+1. B == 1 keeps all."""
    u_coin = np.asarray(u_coin, dtype=np.float32)
    u_pick = np.asarray(u_pick, dtype=np.float32)
    B = u_coin.shape[0]
    i = np.arange(B, dtype=np.int64)
    if B <= 1:
        return np.ones(B, dtype=np.int64), i
    swap = u_coin < np.float32(0.5)
    k = np.floor(u_pick * np.float32(B - 1)).astype(np.int64)
    k = np.minimum(k, B - 2)
    src = np.where(k < i, k, k + 1)
    src = np.where(swap, src, i)
    return (~swap).astype(np.int64), src


def uniforms_from_decisions(swap: np.ndarray, src: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Uniforms that make itm_sample_uniform replay given (swap, src) decisions, e.g. the numpy stream of the CLI."""
    B = swap.shape[0]
    i = np.arange(B)
    k = np.where(src < i, src, src - 1)
    u_coin = np.where(swap, 0.25, 0.75).astype(np.float32)
    u_pick = ((k + 0.5) / max(B - 1, 1)).astype(np.float32)
    u_pick = np.where(swap, u_pick, 0.0).astype(np.float32)
    return u_coin, u_pick


_C1 = np.float32(0.693359375)
_C2 = np.float32(-2.12194440e-4)
_LOG2E = np.float32(1.44269504088896341)
_EXP_POLY = [np.float32(c) for c in (1.9875691500e-4, 1.3981999507e-3, 8.3334519073e-3, 4.1665795894e-2,
                                     1.6666665459e-1, 5.0000001201e-1)]


def det_exp_f32(x: np.ndarray) -> np.ndarray:
    """Bit-reproducible fp32 exp for x <= 0: every step is a single IEEE-754 round-to-nearest fp32 operation
    (no FMA, no library transcendental), mirrored op-for-op by det_exp() in csrc/itm.cu."""
    x = np.maximum(np.asarray(x, dtype=np.float32), np.float32(-80.0))
    n = np.rint(x * _LOG2E).astype(np.float32)
    r = x - n * _C1
    r = r - n * _C2
    p = _EXP_POLY[0]
    for c in _EXP_POLY[1:]:
        p = p * r + c
    y = (p * (r * r) + r) + np.float32(1.0)
    return np.ldexp(y, n.astype(np.int32)).astype(np.float32)


HARD_FIXED_POINT_BITS = 40


def hard_qweights(row: np.ndarray, ref, i: int) -> np.ndarray:
    """Fixed-point weights of one similarity row: q_j = trunc(det_exp(min(S[i,j] - ref, 0)) * 2^40), q_i = 0 (uint64)."""
    x = np.minimum(np.asarray(row, dtype=np.float32) - np.float32(ref), np.float32(0.0)).astype(np.float32)
    q = (det_exp_f32(x) * np.float32(2.0 ** HARD_FIXED_POINT_BITS)).astype(np.uint64)
    q[i] = 0
    return q


def itm_sample_hard(S: np.ndarray, u_coin: np.ndarray, u_pick: np.ndarray, ref=None) -> Tuple[np.ndarray, np.ndarray]:
    """Similarity-weighted hard-negative sampling (extension; ALBEF convention: weights = softmax(S[i,:]) with the
    positive zeroed), made order-independent by fixed-point weights against a FIXED reference `ref` >= max S (the
    product passes exp(logit_scale), the upper bound of the logits and the shift of its softmax statistics; default here:
    the fp32 maximum of S), so that the weights of a row can be summed tile by tile, in any order, without knowing the
    row maximum first:
        q_j = trunc(det_exp(min(S[i,j] - ref, 0)) * 2^40) (uint64), q_i = 0;  total = sum_j q_j
        U = trunc(u_pick * 2^24); target = U*(total>>24) + ((U*(total & (2^24-1))) >> 24)   (= floor(U*total / 2^24))
        src = first j with inclusive-cumsum(q)_j > target.   total == 0 falls back to the uniform rule.
    Rows are swapped iff u_coin < 0.5 exactly as in the uniform rule (mm_late.py:396-399)."""
    S = np.asarray(S, dtype=np.float32)
    u_coin = np.asarray(u_coin, dtype=np.float32)
    u_pick = np.asarray(u_pick, dtype=np.float32)
    B = S.shape[0]
    labels_u, src_u = itm_sample_uniform(u_coin, u_pick)
    if B <= 1:
        return labels_u, src_u
    ref = np.float32(S.max()) if ref is None else np.float32(ref)
    src = src_u.copy()
    for i in np.nonzero(labels_u == 0)[0]:
        q = hard_qweights(S[i], ref, int(i))
        total = int(q.sum(dtype=np.uint64))
        if total == 0:
            continue
        U = int(np.float32(u_pick[i]) * np.float32(16777216.0))
        target = U * (total >> 24) + ((U * (total & 0xFFFFFF)) >> 24)
        c = np.cumsum(q, dtype=np.uint64)
        src[i] = int(np.searchsorted(c, np.uint64(target), side="right"))
    return labels_u, src


def gather_rows(x, src):
    """mm_late.py:403-404 as one gather: out[i] = x[src[i]]."""
    if torch.is_tensor(src):
        return x[src.to(device=x.device, dtype=torch.long)]
    return x[torch.as_tensor(np.asarray(src), dtype=torch.long, device=x.device)]


# ---------------------------------------------------------------------------------------------------- full head


def head_step(inp: Dict[str, torch.Tensor], p: Dict[str, torch.Tensor], *, fusion_name: str = "concat",
              use_itc: bool = True, use_itm: bool = True, beta_itc: float = 0.1, beta_itm: float = 0.1,
              literal_attention: bool = False, relu_masks=None) -> Dict[str, torch.Tensor]:
    """Everything models/mm_late.py does after the encoders return, on given embeddings (forward; use autograd
    on the returned loss for the backward):
      inp: x_t [B,Lt,E], x_v [B,Lv,E], t_pool [B,E], v_pool [B,E], y_soft [B,C], class_w [C] (optional),
           lbl_tim [B] int64 + src_idx [B] int64 (when use_itm), keep [B,E] dropout keep-mask * 1/(1-p) (optional)
      p:   state-dict style names (mm_late.py:59-89): dual_encoder.text_projection.weight, ... .visual_projection.weight,
           dual_encoder.logit_scale, linear_fusion.*, linear_cls.*, linear_tim.*, fc_Q|K|V.*, aspectattention.*, linear_gmu_*.*
    relu_masks (test-only, see _relu): {"main": [B,E], "tim": [B,E]} 0/1 masks used instead of (pre-activation > 0).
    The ITM branch uses x_t[src] — the text encoder is per-sample, so text_model(ids[src]) == text_model(ids)[src]
    (mm_late.py:170-175) — and gets no pools and no dropout (mm_late.py:181-182)."""
    x_t, x_v = inp["x_t"], inp["x_v"]
    out: Dict[str, torch.Tensor] = {}
    T = project(inp["t_pool"], p.get("dual_encoder.text_projection.weight"))
    V = project(inp["v_pool"], p.get("dual_encoder.visual_projection.weight"))
    S = itc_logits(T, V, p["dual_encoder.logit_scale"])                      # mm_late.py:159
    out["logits_per_text"] = S
    pre_main, pre_tim = [], []
    rm = relu_masks or {}
    fused = mm_fusion(fusion_name, x_t, x_v, p, x_v_pool=inp["v_pool"], x_t_pool=inp["t_pool"],
                      literal_attention=literal_attention, relu_mask=rm.get("main"), pre_out=pre_main)   # mm_late.py:160
    out["pre_main"] = pre_main[0]
    out["mm_features"] = fused                                               # :162
    h = fused * inp["keep"] if inp.get("keep") is not None else fused        # :163 (dropout with an injected mask)
    out["out_cls"] = F.linear(h, p["linear_cls.weight"], p["linear_cls.bias"])  # :164
    l_cls = cls_loss_soft(out["out_cls"], inp["y_soft"], inp.get("class_w"))
    l_itc = clip_loss(S) if use_itc else None
    l_itm = None
    if use_itm:
        x_t_tim = gather_rows(x_t, inp["src_idx"])                           # :170-175 via permutation equivariance
        fused_tim = mm_fusion(fusion_name, x_t_tim, x_v, p, literal_attention=literal_attention,
                              relu_mask=rm.get("tim"), pre_out=pre_tim)                          # :181 (no pools)
        out["pre_tim"] = pre_tim[0]
        out["out_tim"] = F.linear(fused_tim, p["linear_tim.weight"], p["linear_tim.bias"])        # :182
        l_itm = itm_loss(out["out_tim"], inp["lbl_tim"])
    out["loss_cls"], out["loss_itc"], out["loss_itm"] = l_cls, l_itc, l_itm
    out["loss"] = loss_mix(l_cls, l_itc, l_itm, use_itc, use_itm, beta_itc, beta_itm)
    return out


def init_params(num_labels: int, E: int = 768, P: Optional[int] = 512, seed: int = 40, dtype=torch.float32):
    """nn.Linear-style U(-1/sqrt(in), 1/sqrt(in)) init of the head layers (mm_late.py:71-89) and the HF projection
    layers, drawn from numpy's legacy MT19937 stream so fixtures regenerate identically on any torch version."""
    rs = np.random.RandomState(seed)

    def lin(o, i, bias=True):
        bound = 1.0 / math.sqrt(i)
        w = torch.from_numpy(rs.uniform(-bound, bound, size=(o, i))).to(dtype)
        b = torch.from_numpy(rs.uniform(-bound, bound, size=(o,))).to(dtype) if bias else None
        return w, b

    p = {}
    if P is not None:
        p["dual_encoder.visual_projection.weight"], _ = lin(P, E, False)
        p["dual_encoder.text_projection.weight"], _ = lin(P, E, False)
    p["dual_encoder.logit_scale"] = torch.tensor(2.6592, dtype=dtype)
    for name, (o, i) in (("fc_Q", (E, E)), ("fc_K", (E, E)), ("fc_V", (E, E)), ("aspectattention", (1, E)),
                         ("linear_fusion", (E, 2 * E)), ("linear_cls", (num_labels, E)), ("linear_tim", (2, E)),
                         ("linear_iadds", (2, E)), ("linear_gmu_t", (2 * E, E)), ("linear_gmu_v", (2 * E, E))):
        p[name + ".weight"], p[name + ".bias"] = lin(o, i)
    return p


# ------------------------------------------------------------------------------------------------ eval bookkeeping (§8 f-4)
def eval_batch(output: torch.Tensor, label: torch.Tensor):
    """models/mm_late.py:596-608 (single-label branch): pred = argmax(softmax(output)), target = argmax(label) of the float
    one-hot labels, accuracy of this batch in percent.  Returns (pred int64 [B], target int64 [B], accuracy float)."""
    pred = torch.argmax(torch.softmax(output.double(), dim=1), dim=1)
    target = torch.argmax(label, dim=1)
    return pred, target, float((pred == target).double().mean()) * 100.0


def eval_epoch(outputs, labels, losses) -> Dict:
    """models/mm_late.py:544-636: the per-batch loop of MMLate_Model.eval on given per-batch logits / float one-hot labels /
    batch losses: loss = mean of the batch losses (:594,615), accuracy = mean of the per-batch accuracies (:607-608,616 — a short
    last batch weighs as much as a full one), predictions / labels concatenated in order (:610-611,623,628)."""
    preds, tgts, accs = [], [], []
    for out, lab in zip(outputs, labels):
        p, t, a = eval_batch(out, lab)
        preds.append(p); tgts.append(t); accs.append(a)
    return {"loss": float(np.mean([float(l) for l in losses])), "accuracy": float(np.mean(accs)),
            "predictions": torch.cat(preds), "labels": torch.cat(tgts)}


def confusion_matrix(pred: np.ndarray, target: np.ndarray, C: int) -> np.ndarray:
    """conf[t, p] = number of samples with target t predicted as p (the statistic every torchmetrics score below reduces)."""
    conf = np.zeros((C, C), dtype=np.int64)
    np.add.at(conf, (np.asarray(target, dtype=np.int64), np.asarray(pred, dtype=np.int64)), 1)
    return conf


def metrics_from_confusion(conf: np.ndarray) -> Dict[str, float]:
    """models/utils.py:294-325 compute_metrics, single-label branch: F1 / precision / recall, `weighted` and `macro`, as
    torchmetrics==0.11.0 (timrel-env.yml:120) computes them for task="multiclass".  torchmetrics is a third-party dependency
    that is NOT in the reference tree and NOT installed here, so this restates its published algorithm
    (functional/classification/{f_beta,precision_recall}.py: per-class tp/fp/fn from the confusion matrix, _safe_divide -> 0
    where the denominator is 0, _adjust_weights_safe_divide: `weighted` weighs classes by support tp+fn; `macro` averages the
    classes with tp+fp+fn > 0).  PARITY UNPINNED against torchmetrics itself; tests pin it against scikit-learn where the
    semantics coincide (every class present)."""
    conf = np.asarray(conf, dtype=np.float64)
    tp = np.diag(conf)
    fn = conf.sum(axis=1) - tp
    fp = conf.sum(axis=0) - tp

    def sdiv(a, b):
        out = np.zeros_like(a)
        np.divide(a, b, out=out, where=b != 0)
        return out

    prec, rec, f1 = sdiv(tp, tp + fp), sdiv(tp, tp + fn), sdiv(2 * tp, 2 * tp + fp + fn)
    w_sup = tp + fn
    w_mac = ((tp + fp + fn) > 0).astype(np.float64)

    def avg(score, w):
        return float((score * w).sum() / w.sum()) if w.sum() > 0 else 0.0

    return {"f1_weighted": avg(f1, w_sup), "f1_macro": avg(f1, w_mac), "precision_weighted": avg(prec, w_sup),
            "precision_macro": avg(prec, w_mac), "recall_weighted": avg(rec, w_sup), "recall_macro": avg(rec, w_mac)}


def compute_metrics(res: Dict, num_classes: int) -> Dict[str, list]:
    """models/utils.py:294-325: {"metric": [six names, "loss"], "result": [...]} from an eval() result dict."""
    pred = torch.as_tensor(res["predictions"]).cpu().numpy()
    tgt = torch.as_tensor(res["labels"]).cpu().numpy()
    m = metrics_from_confusion(confusion_matrix(pred, tgt, num_classes))
    m["loss"] = res["loss"]
    return {"metric": list(m.keys()), "result": list(m.values())}
