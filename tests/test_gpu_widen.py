"""GPU: the rows SURVEY.md §8 marks "next" — f-2 device-side prepare_itm_inputs, f-3 the mm_early auxiliary-loss tail,
f-4 evaluation bookkeeping on the device — each against the CPU oracle on the same inputs.  Integer work (sampled indices,
gathered rows, predictions, confusion counts) is bit-exact; floating point within the 1e-3 of BASELINE.json (the metric
scores, being ratios of exact counts, to 1e-6)."""
import os
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import restatement as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


# ------------------------------------------------------------------------------------------------ f-4 eval bookkeeping
@pytest.mark.parametrize("C,sizes,int_labels", [(4, [8, 8, 8, 3], False), (2, [256, 256, 100], False), (3, [1], False),
                                                (4, [64, 64], True), (20, [500, 37], False), (4, [5000], False)])
def test_eval_accumulator_matches_reference_eval_loop(C, sizes, int_labels):
    from tic_b200.eval import EvalAccumulator, METRIC_NAMES
    g = torch.Generator().manual_seed(C * 100 + sum(sizes))
    acc = EvalAccumulator(C, sum(sizes), device=DEV)
    preds, tgts, accs, losses = [], [], [], []
    for B in sizes:
        out = torch.randn(B, C, generator=g)
        out[:: 7, 0] = out[:: 7, 1]                      # exact ties: first index wins (torch.argmax)
        y = torch.randint(0, max(C - 1, 1), (B,), generator=g)   # the last class never occurs as a target
        label = torch.eye(C)[y]
        loss = torch.rand((), generator=g)
        p, t, a = R.eval_batch(out, label)
        preds.append(p); tgts.append(t); accs.append(a); losses.append(float(loss))
        acc.update(out.to(DEV), y.to(DEV) if int_labels else label.to(DEV), loss=loss.to(DEV),
                   data_id=torch.arange(B, device=DEV))
    res = acc.result()
    pred_ref, tgt_ref = torch.cat(preds), torch.cat(tgts)
    assert torch.equal(res["predictions"].cpu(), pred_ref) and torch.equal(res["labels"].cpu(), tgt_ref)
    conf_ref = R.confusion_matrix(pred_ref.numpy(), tgt_ref.numpy(), C)
    assert np.array_equal(acc.confusion.cpu().numpy(), conf_ref)
    assert abs(res["loss"] - float(np.mean(losses))) < 1e-5 * max(1.0, abs(float(np.mean(losses))))
    assert abs(res["accuracy"] - float(np.mean(accs))) < 1e-3          # fp32 sum of per-batch percentages
    m_ref = R.metrics_from_confusion(conf_ref)
    got = acc.metrics_device().cpu().tolist()
    for name, v in zip(METRIC_NAMES, got):
        assert abs(v - m_ref[name]) < 1e-6, name
    assert res["data_id"].shape[0] == sum(sizes)


def test_compute_metrics_mirror_matches_oracle_and_sklearn():
    from sklearn.metrics import f1_score, precision_score, recall_score
    from tic_b200.utils import compute_metrics
    rs = np.random.RandomState(3)
    t = rs.randint(0, 4, 1000)
    p = np.where(rs.rand(1000) < 0.6, t, rs.randint(0, 4, 1000))
    res = {"predictions": torch.tensor(p, device=DEV), "labels": torch.tensor(t, device=DEV), "loss": 0.25}
    got = compute_metrics(res, 4)
    ref = R.compute_metrics({"predictions": torch.tensor(p), "labels": torch.tensor(t), "loss": 0.25}, 4)
    assert got["metric"] == ref["metric"]
    for a, b in zip(got["result"], ref["result"]):
        assert abs(a - b) < 1e-6
    d = dict(zip(got["metric"], got["result"]))
    assert abs(d["f1_macro"] - f1_score(t, p, average="macro")) < 1e-6
    assert abs(d["precision_weighted"] - precision_score(t, p, average="weighted")) < 1e-6
    assert abs(d["recall_macro"] - recall_score(t, p, average="macro")) < 1e-6


# ------------------------------------------------------------------------------------------------ f-2 device-side ITM inputs
@pytest.mark.parametrize("B", [1, 2, 8, 257, 4096])
def test_prepare_itm_inputs_device_rng_bit_exact_given_the_uniforms(B):
    from tic_b200.mm_late import MMLate_Model
    self = types.SimpleNamespace()
    ids = torch.randint(5, 30000, (B, 128), device=DEV)
    mask = (torch.rand(B, 128, device=DEV) < 0.9).to(torch.int64)
    g = torch.Generator(device=DEV).manual_seed(1234 + B)
    tim_ids, tim_mask, lbl, src = MMLate_Model.prepare_itm_inputs(self, ids, mask, return_src=True, rng="device", generator=g)
    g2 = torch.Generator(device=DEV).manual_seed(1234 + B)
    u = torch.rand(2, B, device=DEV, generator=g2).cpu().numpy()
    lbl_ref, src_ref = R.itm_sample_uniform(u[0], u[1])
    assert np.array_equal(lbl.cpu().numpy(), lbl_ref) and np.array_equal(src.cpu().numpy(), src_ref.astype(np.int32))
    assert torch.equal(tim_ids, ids[src.long()]) and torch.equal(tim_mask, mask[src.long()])
    assert tim_ids.data_ptr() != ids.data_ptr() and tim_mask.data_ptr() != mask.data_ptr()     # fresh clones (:391-392)
    if B > 1:
        assert 0.3 < float((lbl == 0).float().mean()) < 0.7 or B < 64
        assert bool(((src.long() != torch.arange(B, device=DEV)) == (lbl == 0)).all())         # swapped rows never pick themselves


# ------------------------------------------------------------------------------------------------ f-3 mm_early tail
def _oracle_logits(T, V, ls):
    Tr = T.double().requires_grad_(True)
    Vr = V.double().requires_grad_(True)
    l = torch.tensor(float(ls), dtype=torch.float64, requires_grad=True)
    return Tr, Vr, l, R.itc_logits(Tr, Vr, l)


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("B,d", [(8, 768), (100, 768), (256, 768), (33, 64)])
def test_mm_early_get_logits_per_text_and_gradients(B, d):
    from tic_b200.mm_early import get_logits_per_text
    from tic_b200.utils import clip_loss
    g = torch.Generator().manual_seed(B * 3 + d)
    T32 = torch.randn(B, d, generator=g)
    V32 = torch.randn(B, d, generator=g) + 0.3 * T32
    W = torch.randn(B, B, generator=g)
    T = T32.to(DEV).requires_grad_(True)
    V = V32.to(DEV).requires_grad_(True)
    ls = torch.tensor(2.6592, device=DEV, requires_grad=True)
    S = get_logits_per_text(T, V, ls)
    # an arbitrary downstream loss: the reference's clip_loss plus a linear functional (exercises a generic dS)
    loss = clip_loss(S) + (S * W.to(DEV)).sum() * 1e-3
    loss.backward()
    Tr, Vr, l, Sr = _oracle_logits(T32, V32, 2.6592)
    loss_r = R.clip_loss(Sr) + (Sr * W.double()).sum() * 1e-3
    loss_r.backward()
    assert _rel(S, Sr) < 1e-3
    assert abs(float(loss) - float(loss_r)) / abs(float(loss_r)) < 1e-3
    assert _rel(T.grad, Tr.grad) < 1e-3 and _rel(V.grad, Vr.grad) < 1e-3
    assert abs(float(ls.grad) - float(l.grad)) / max(abs(float(l.grad)), 1e-6) < 1e-3


@pytest.mark.parametrize("B,d", [(8, 768), (33, 64), (48, 256)])
def test_mm_early_tail_vs_reference_golden(golden_dir, B, d):
    """Against values and gradients recorded from the UNMODIFIED reference (ViLT.get_logits_per_text + utils.clip_loss,
    tests/golden/mm_early_tail.npz by oracle/make_golden.py), same loss as the fixture: clip_loss(S) + 1e-3 * <S, W>."""
    from tic_b200.mm_early import get_logits_per_text, itc_loss
    from tic_b200.utils import clip_loss
    gd = dict(np.load(os.path.join(golden_dir, "mm_early_tail.npz")))
    key = "b%d_d%d_" % (B, d)
    T = torch.tensor(gd[key + "T"], device=DEV, requires_grad=True)
    V = torch.tensor(gd[key + "V"], device=DEV, requires_grad=True)
    ls = torch.tensor(2.6592, device=DEV, requires_grad=True)
    S = get_logits_per_text(T, V, ls)
    closs = clip_loss(S)
    (closs + 1e-3 * (S * torch.tensor(gd[key + "W"], device=DEV)).sum()).backward()
    assert _rel(S, torch.tensor(gd[key + "S"])) < 1e-3
    assert abs(float(closs) - float(gd[key + "clip_loss"])) / float(gd[key + "clip_loss"]) < 1e-3
    assert _rel(T.grad, torch.tensor(gd[key + "dT"])) < 1e-3 and _rel(V.grad, torch.tensor(gd[key + "dV"])) < 1e-3
    assert abs(float(ls.grad) - float(gd[key + "dls"])) / abs(float(gd[key + "dls"])) < 1e-3
    fused = itc_loss(T.detach(), V.detach(), ls.detach())
    assert abs(float(fused) - float(gd[key + "clip_loss"])) / float(gd[key + "clip_loss"]) < 1e-3


@pytest.mark.parametrize("B,d", [(8, 768), (256, 768), (1000, 256), (4096, 768)])
def test_mm_early_fused_itc_loss(B, d):
    from tic_b200.mm_early import itc_loss
    g = torch.Generator().manual_seed(B + d)
    T32 = torch.randn(B, d, generator=g)
    V32 = torch.randn(B, d, generator=g) + 0.3 * T32
    if B >= 4096:      # from 4096 negatives on the embeddings are consumed as single bf16 (the residual K-segments are skipped where
        T32, V32 = T32.to(torch.bfloat16).float(), V32.to(torch.bfloat16).float()   # tensor time matters): bf16-representable inputs
    T = T32.to(DEV).requires_grad_(True)
    V = V32.to(DEV).requires_grad_(True)
    ls = torch.tensor(2.6592, device=DEV, requires_grad=True)
    loss = itc_loss(T, V, ls)
    (0.1 * loss).backward()        # beta_itc = 0.1 upstream
    Tr, Vr, l, Sr = _oracle_logits(T32, V32, 2.6592)
    loss_r = R.clip_loss(Sr)
    (0.1 * loss_r).backward()
    assert abs(float(loss) - float(loss_r)) / abs(float(loss_r)) < 1e-3
    assert _rel(T.grad, Tr.grad) < 1e-3 and _rel(V.grad, Vr.grad) < 1e-3
    assert abs(float(ls.grad) - float(l.grad)) / max(abs(float(l.grad)), 1e-6) < 1e-3


@pytest.mark.parametrize("seed,B", [(40, 8), (30, 16)])
def test_mm_early_prepare_itm_inputs_replays_the_numpy_stream(golden_dir, seed, B):
    """Same stream contract as the late-fusion twin (SURVEY §8c golden vectors): seed 40, B=8 and seed 30, B=16."""
    from tic_b200.mm_early import prepare_itm_inputs
    gold = dict(np.load(os.path.join(golden_dir, "itm_stream.npz")))
    ids = torch.arange(B * 6, device=DEV).view(B, 6)
    mask = torch.ones_like(ids)
    tt = (torch.arange(B * 6, device=DEV).view(B, 6) % 2)
    np.random.seed(seed)
    tim_ids, tim_mask, tim_tt, lbl = prepare_itm_inputs(ids, mask, tt)
    np.random.seed(seed)
    swap_ref, src_ref = R.itm_decisions_from_stream(B, np.random)
    lbl_ref = (~swap_ref).astype(np.int64)
    assert np.array_equal(gold["s%d_b%d_lbl" % (seed, B)], lbl_ref) and np.array_equal(gold["s%d_b%d_src" % (seed, B)], src_ref)
    assert np.array_equal(lbl.cpu().numpy(), lbl_ref)
    s = torch.as_tensor(np.asarray(src_ref), device=DEV).long()
    assert torch.equal(tim_ids, ids[s]) and torch.equal(tim_mask, mask[s]) and torch.equal(tim_tt, tt[s])
    # device rng: same rule on device-drawn uniforms
    g = torch.Generator(device=DEV).manual_seed(7)
    a, b, c, l2 = prepare_itm_inputs(ids, mask, tt, rng="device", generator=g)
    g2 = torch.Generator(device=DEV).manual_seed(7)
    u = torch.rand(2, B, device=DEV, generator=g2).cpu().numpy()
    l_ref, s_ref = R.itm_sample_uniform(u[0], u[1])
    assert np.array_equal(l2.cpu().numpy(), l_ref)
    assert torch.equal(a, ids[torch.as_tensor(s_ref, device=DEV).long()])


def test_mm_early_aux_loss_mix_matches_reference_formula():
    from tic_b200.mm_early import aux_loss_mix, get_logits_per_text
    g = torch.Generator().manual_seed(5)
    T32, V32 = torch.randn(16, 768, generator=g), torch.randn(16, 768, generator=g)
    S = get_logits_per_text(T32.to(DEV), V32.to(DEV), torch.tensor(2.6592, device=DEV))
    l_cls, l_itm = torch.tensor(1.25, device=DEV), torch.tensor(0.7, device=DEV)
    Sr = R.itc_logits(T32.double(), V32.double(), torch.tensor(2.6592, dtype=torch.float64))
    ref = R.loss_mix(torch.tensor(1.25).double(), R.clip_loss(Sr), torch.tensor(0.7).double(), True, True, 0.1, 0.1)
    got = aux_loss_mix(l_cls, S, l_itm, 0.1, 0.1)
    assert abs(float(got) - float(ref)) / float(ref) < 1e-3
    ref2 = R.loss_mix(torch.tensor(1.25).double(), None, torch.tensor(0.7).double(), False, True, 0.1, 0.1)
    assert abs(float(aux_loss_mix(l_cls, None, l_itm, 0.1, 0.1)) - float(ref2)) < 1e-6


# ------------------------------------------------------------------------------------------------ trainer wrapper: eval & co
class _Tower(nn.Module):
    def __init__(self, hidden, pool):
        super().__init__()
        self.h, self.p = nn.Parameter(hidden.clone()), nn.Parameter(pool.clone())
        self.rows = None

    def forward(self, **kw):
        return types.SimpleNamespace(last_hidden_state=self.h[self.rows], pooler_output=self.p[self.rows])


class _StubDualEncoder(nn.Module):
    def __init__(self, x_t, t_pool, x_v, v_pool, P=512, E=768):
        super().__init__()
        self.vision_model, self.text_model = _Tower(x_v, v_pool), _Tower(x_t, t_pool)
        self.visual_projection = nn.Linear(E, P, bias=False)
        self.text_projection = nn.Linear(E, P, bias=False)
        self.logit_scale = nn.Parameter(torch.tensor(2.6592))


def test_mmlate_model_eval_predictions_features_on_device():
    from tic_b200.mm_late import MM_Model, MMLate_Model
    N, C, bs = 20, 4, 8
    g = torch.Generator().manual_seed(11)
    bf = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).float()   # noqa: E731
    x_t, t_pool, x_v, v_pool = bf(N, 2, 768), torch.tanh(bf(N, 768)), bf(N, 2, 768), torch.tanh(bf(N, 768))
    de = _StubDualEncoder(x_t, t_pool, x_v, v_pool)
    model = MM_Model(C, "bert", "vit", 0.05, fusion_name="concat", dual_encoder=de)
    p = R.init_params(C, seed=40)
    sd = model.state_dict()
    for k, v in p.items():
        sd[k].copy_(v)
    cfg = types.SimpleNamespace(batch_size=bs, num_labels=C, use_clip_loss=True, beta_itc=0.1, use_tim_loss=True, beta_itm=0.1,
                                use_iadds_loss=False, beta_iadds=0.0, use_loss_correction=False, max_length=128, dropout=0.05)
    wrap = MMLate_Model(cfg, "bert", "vit", "concat", model=model, device=DEV, itm_rng="device")
    y = torch.randint(0, C, (N,), generator=g)
    labels = torch.eye(C)[y]

    class Loader(list):
        dataset = list(range(N))
    batches = Loader()
    for s in range(0, N, bs):
        rows = torch.arange(s, min(s + bs, N))
        batches.append({"input_ids": torch.randint(5, 999, (len(rows), 1, 16), generator=g),
                        "attention_mask": torch.ones(len(rows), 1, 16, dtype=torch.int64),
                        "pixel_values": torch.zeros(len(rows), 1, 3, 8, 8), "labels": labels[rows], "data_id": rows.clone(),
                        "_rows": rows})
    orig = wrap._batch

    def batch_hook(b):        # route the stub towers to this batch's rows
        de.vision_model.rows = de.text_model.rows = b["_rows"].to(DEV)
        return orig(b)
    wrap._batch = batch_hook
    loss_fn, tim_loss_fn = nn.CrossEntropyLoss(weight=torch.ones(C, device=DEV)), nn.CrossEntropyLoss()
    res = wrap.eval(batches, loss_fn, tim_loss_fn=tim_loss_fn)
    assert res["predictions"].shape == (N,) and torch.equal(res["labels"].cpu(), y) and torch.equal(res["data_id"].cpu(), torch.arange(N))
    # oracle: the main-branch logits do not depend on the ITM sampling
    pd = {k: (v.to(torch.bfloat16) if (v.dim() == 2 and v.numel() > 8 * 768 and not k.startswith(("linear_cls", "linear_tim"))) else v).double()
          for k, v in p.items()}
    ref_pred = []
    for s in range(0, N, bs):
        r = slice(s, min(s + bs, N))
        inp = {"x_t": x_t[r].double(), "x_v": x_v[r].double(), "t_pool": t_pool[r].double(), "v_pool": v_pool[r].double(),
               "y_soft": labels[r].double()}
        out = R.head_step(inp, pd, fusion_name="concat", use_itc=True, use_itm=False)
        ref_pred.append(torch.argmax(out["out_cls"], dim=1))
    ref_pred = torch.cat(ref_pred)
    assert torch.equal(res["predictions"].cpu(), ref_pred)
    conf = wrap.last_eval.confusion.cpu().numpy()
    assert np.array_equal(conf, R.confusion_matrix(ref_pred.numpy(), y.numpy(), C)) and np.isfinite(res["loss"])
    pr = wrap.compute_predictions(batches)
    assert torch.equal(pr["predictions"].cpu(), ref_pred) and torch.equal(pr["data_id"].cpu(), torch.arange(N))
    feats, ys = wrap.extract_features(batches)
    assert feats.shape == (N, 768) and torch.equal(ys.cpu(), y)
    from tic_b200.utils import compute_metrics
    m = compute_metrics(res, C)
    mr = R.compute_metrics({"predictions": ref_pred, "labels": y, "loss": res["loss"]}, C)
    for a, b in zip(m["result"], mr["result"]):
        assert abs(a - b) < 1e-6
