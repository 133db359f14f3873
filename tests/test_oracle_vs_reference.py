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
