"""CPU, build container only: the oracle restatement against the LIVE unmodified reference imported from
/root/reference under oracle/ref_shims.py (skipped where the tree is absent, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import ref_shims
from oracle import restatement as R

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ns():
    return ref_shims.load_reference(small_encoders=True)


@pytest.mark.parametrize("B", [2, 5, 64])
def test_clip_loss_live(ns, B):
    S = torch.randn(B, B, generator=torch.Generator().manual_seed(B)) * 4
    assert torch.allclose(ns.utils.clip_loss(S), R.clip_loss(S), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("seed,B", [(0, 3), (11, 8), (123, 100), (40, 1000)])
def test_prepare_itm_inputs_live(ns, seed, B):
    ids = torch.arange(B * 4).view(B, 4)
    mask = (ids % 3 != 0).long()
    np.random.seed(seed)
    a, b, l = ref_shims.ref_prepare_itm_inputs(ns, ids, mask)
    a2, b2, l2 = R.prepare_itm_inputs_stream(ids, mask, np.random.RandomState(seed))
    assert torch.equal(a, a2) and torch.equal(b, b2) and torch.equal(l, l2)


@pytest.mark.parametrize("fusion", ["concat", "attention", "gmu", "aspect-att"])
def test_mm_fusion_live(ns, fusion, capsys):
    torch.manual_seed(0)
    m = ns.mm_late.MM_Model(3, "bert", "vit", 0.0, fusion_name=fusion).eval()
    p = {k: v.detach() for k, v in m.named_parameters() if "text_model" not in k and "vision_model" not in k}
    g = torch.Generator().manual_seed(1)
    x_t, x_v = torch.randn(5, 9, 768, generator=g), torch.randn(5, 7, 768, generator=g)
    tp, vp = torch.randn(5, 768, generator=g), torch.randn(5, 768, generator=g)
    with torch.no_grad():
        ref = m.mm_fusion(x_t, x_v, x_v_pool=vp, x_t_pool=tp)
        got = R.mm_fusion(fusion, x_t, x_v, p, x_v_pool=vp, x_t_pool=tp)
        lit = R.mm_fusion(fusion, x_t, x_v, p, x_v_pool=vp, x_t_pool=tp, literal_attention=True)
    assert torch.allclose(ref, got, rtol=1e-4, atol=2e-6)
    assert torch.allclose(ref, lit, rtol=1e-5, atol=1e-6)


def test_itc_logits_live_hf(ns):
    """HF VisionTextDualEncoderModel.forward tail (the third-party arithmetic) == restatement on its own embeds."""
    torch.manual_seed(0)
    m = ns.mm_late.MM_Model(3, "bert", "vit", 0.0).eval()
    de = m.dual_encoder
    ids = torch.randint(5, 1000, (3, 7))
    pix = torch.randn(3, 3, de.config.vision_config.image_size, de.config.vision_config.image_size)
    with torch.no_grad():
        out = de(input_ids=ids, attention_mask=torch.ones_like(ids), pixel_values=pix, return_loss=True)
        S = R.itc_logits(R.project(out.text_model_output.pooler_output, de.text_projection.weight),
                         R.project(out.vision_model_output.pooler_output, de.visual_projection.weight), de.logit_scale)
    assert torch.allclose(out.logits_per_text, S, rtol=1e-5, atol=1e-6)
    assert torch.allclose(out.loss, R.clip_loss(S), rtol=1e-5)


# ---------------------------------------------------------------------------------------------------- mm_early tail (§8 f-3)
@pytest.fixture(scope="module")
def ns_early():
    return ref_shims.load_reference_early()


@pytest.mark.parametrize("B,d", [(2, 768), (17, 768), (64, 128)])
def test_mm_early_get_logits_per_text_live(ns_early, B, d):
    """ViLT / Lxmert.get_logits_per_text (mm_early.py:96-103, 165-172) == the restatement, values and autograd gradients."""
    g = torch.Generator().manual_seed(B + d)
    T0, V0 = torch.randn(B, d, generator=g), torch.randn(B, d, generator=g)
    outs = []
    for fn in (lambda T, V, ls: ref_shims.ref_early_logits(ns_early, T, V, ls), R.itc_logits):
        T, V = T0.clone().requires_grad_(True), V0.clone().requires_grad_(True)
        ls = torch.tensor(2.6592, requires_grad=True)
        S = fn(T, V, ls)
        ns_early.utils.clip_loss(S).backward()
        outs.append((S.detach(), T.grad, V.grad, ls.grad))
    for a, b in zip(*outs):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
    lx = ns_early.mm_early.Lxmert.get_logits_per_text
    import types as _t
    S2 = lx(_t.SimpleNamespace(logit_scale=torch.tensor(2.6592)), T0, V0)
    assert torch.allclose(S2, outs[0][0], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("seed,B", [(0, 3), (11, 8), (40, 300)])
def test_mm_early_prepare_itm_inputs_live(ns_early, seed, B):
    ids = torch.arange(B * 4).view(B, 4)
    mask, tt = (ids % 3 != 0).long(), ids % 2
    np.random.seed(seed)
    a, b, c, l = ref_shims.ref_early_prepare_itm_inputs(ns_early, ids, mask, tt)
    swap, src = R.itm_decisions_from_stream(B, np.random.RandomState(seed))
    assert torch.equal(l, torch.from_numpy((~swap).astype(np.int64)))
    assert torch.equal(a, R.gather_rows(ids, src)) and torch.equal(b, R.gather_rows(mask, src)) and torch.equal(c, R.gather_rows(tt, src))
    assert a.data_ptr() != ids.data_ptr()        # fresh clones (mm_early.py:265-267)


# ---------------------------------------------------------------------------------------------------- eval bookkeeping (§8 f-4)
@pytest.mark.parametrize("C,sizes", [(4, [8, 8, 3]), (2, [5]), (3, [16, 16, 16, 1])])
def test_eval_loop_live(ns, C, sizes, capsys):
    """MMLate_Model.eval (mm_late.py:534-638), run unbound on replayed logits, == R.eval_epoch: predictions, labels, mean batch loss."""
    g = torch.Generator().manual_seed(C + sum(sizes))
    outs = [torch.randn(B, C, generator=g) for B in sizes]
    labs = [torch.eye(C)[torch.randint(0, C, (B,), generator=g)] for B in sizes]
    loss_fn = torch.nn.CrossEntropyLoss(weight=torch.rand(C, generator=g) + 0.5)
    ref = ref_shims.ref_eval(ns, outs, labs, loss_fn)
    mine = R.eval_epoch(outs, labs, [loss_fn(o, l) for o, l in zip(outs, labs)])
    assert torch.equal(ref["predictions"], mine["predictions"]) and torch.equal(ref["labels"], mine["labels"])
    assert abs(float(ref["loss"]) - mine["loss"]) < 1e-6
    assert torch.equal(ref["data_id"], torch.arange(sum(sizes)))
    printed = capsys.readouterr().out          # the reference prints "loss: … acc: …" (mm_late.py:618): pin the accuracy too
    acc = float(printed.split("acc:")[1].split()[0])
    assert abs(acc - mine["accuracy"]) < 1e-3
