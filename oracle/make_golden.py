"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference under oracle/ref_shims.py) in this container.  The fixtures travel to the GPU box; the reference
does not.  Re-run with:  python -m oracle.make_golden

What is recorded
  itm_stream.npz      MMLate_Model.prepare_itm_inputs (mm_late.py:389-414) on the global numpy stream for
                      (seed, B) in {(40,8), (30,16), (123,256), (0,2), (7,1)}: labels, gathered ids/mask, source rows.
  clip_loss.npz       utils.clip_loss (utils.py:225-231) values + autograd gradients on seeded matrices.
  mm_early_tail.npz   the encoder-free tail of the early-fusion models: ViLT.get_logits_per_text
                      (mm_early.py:96-103) + utils.clip_loss with autograd gradients w.r.t. both embeddings and logit_scale,
                      and MMEarly_Model.prepare_itm_inputs (mm_early.py:262-293) on the numpy stream.
  head_<fusion>.npz   MM_Model.forward (mm_late.py:148-193) end to end through tiny random-init encoders, eval mode,
                      plus the reference loss code (run_mm_late.py:85,97; mm_late.py:473-487) and autograd:
                      captured encoder outputs (= the head's inputs), head outputs, losses, gradients of every head
                      parameter (full for small ones, checksums + slices for the large ones) and of the encoder outputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from oracle import restatement as R  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
HEAD_PARAM_NAMES = ["dual_encoder.logit_scale", "dual_encoder.visual_projection.weight",
                    "dual_encoder.text_projection.weight", "fc_Q.weight", "fc_Q.bias", "fc_K.weight", "fc_K.bias",
                    "fc_V.weight", "fc_V.bias", "aspectattention.weight", "aspectattention.bias", "linear_fusion.weight",
                    "linear_fusion.bias", "linear_cls.weight", "linear_cls.bias", "linear_tim.weight", "linear_tim.bias",
                    "linear_gmu_t.weight", "linear_gmu_t.bias", "linear_gmu_v.weight", "linear_gmu_v.bias"]


def grad_summary(g: torch.Tensor):
    """Large gradients are stored as (sum, abs-sum, l2, leading 4x8 block) instead of the full tensor."""
    g = g.detach().double()
    flat = g.reshape(g.shape[0], -1) if g.dim() > 1 else g.reshape(1, -1)
    return np.array([g.sum().item(), g.abs().sum().item(), g.norm().item()]), flat[:4, :8].numpy().copy()


def gen_itm_stream(ns):
    out = {}
    for seed, B in ((40, 8), (30, 16), (123, 256), (0, 2), (7, 1)):
        L = 6
        ids = (torch.arange(B * L).view(B, L) * 7 + 3) % 1000
        mask = ((torch.arange(B * L).view(B, L) % 5) != 0).long()
        np.random.seed(seed)
        tim_ids, tim_mask, lbl = ref_shims.ref_prepare_itm_inputs(ns, ids, mask)
        src = (tim_ids[:, 0] - 3) % 1000  # ids[:,0] = (7*L*row + 3) % 1000 is unique for B*L*7 < 1000 only; recover robustly:
        src = torch.tensor([int((ids == tim_ids[i]).all(dim=1).nonzero()[0]) for i in range(B)])
        key = "s%d_b%d" % (seed, B)
        out[key + "_ids"] = ids.numpy()
        out[key + "_mask"] = mask.numpy()
        out[key + "_tim_ids"] = tim_ids.numpy()
        out[key + "_tim_mask"] = tim_mask.numpy()
        out[key + "_lbl"] = lbl.numpy()
        out[key + "_src"] = src.numpy()
    np.savez_compressed(os.path.join(GOLD, "itm_stream.npz"), **out)
    print("itm_stream:", out["s40_b8_lbl"].tolist(), out["s40_b8_src"].tolist())


def gen_clip_loss(ns):
    out = {}
    rs = np.random.RandomState(40)
    for B in (1, 2, 8, 33, 128):
        S = torch.tensor(rs.normal(0, 3.0, size=(B, B)), dtype=torch.float32, requires_grad=True)
        loss = ns.utils.clip_loss(S)
        loss.backward()
        out["b%d_S" % B] = S.detach().numpy()
        out["b%d_loss" % B] = loss.detach().numpy()
        out["b%d_dS" % B] = S.grad.numpy()
    # structural known answers
    for B, c in ((8, 5.0), (16, 0.0)):
        S = c * torch.eye(B)
        out["eye%d_c%g" % (B, c)] = ns.utils.clip_loss(S).numpy()
    np.savez_compressed(os.path.join(GOLD, "clip_loss.npz"), **out)
    print("clip_loss b8:", out["b8_loss"])


def gen_head(ns, fusion, use_itm, B=4, C=4, Lt=12, seed=40):
    torch.manual_seed(seed)
    model = ns.mm_late.MM_Model(C, "bert", "vit", 0.05, fusion_name=fusion)
    model.eval()  # dropout off (both the head's and the encoders')
    params = R.init_params(C, seed=seed)
    sd = model.state_dict()
    with torch.no_grad():
        for k, v in params.items():
            assert k in sd and tuple(sd[k].shape) == tuple(v.shape), (k, sd[k].shape, v.shape)
            dict(model.named_parameters())[k].copy_(v)
    rs = np.random.RandomState(seed + 1)
    vcfg = model.dual_encoder.config.vision_config
    ids = torch.from_numpy(rs.randint(5, 2000, size=(B, Lt))).long()
    mask = torch.ones(B, Lt, dtype=torch.long)
    mask[1, Lt - 3:] = 0
    pixels = torch.from_numpy(rs.normal(0, 1, size=(B, 3, vcfg.image_size, vcfg.image_size))).float()
    y = rs.randint(0, C, size=B)
    y_soft = torch.eye(C)[torch.from_numpy(y)]
    class_w = torch.from_numpy(rs.uniform(0.5, 2.0, size=C)).float()

    # capture encoder outputs (the head's inputs) and their gradients
    cap = {"text": [], "vision": []}

    def text_hook(_m, _inp, outp):
        # Cut the graph at the encoder boundary: the head's inputs become leaves, so their .grad is exactly the
        # head's gradient (otherwise last_hidden_state.grad also carries the BERT pooler's back-propagation of t_pool).
        outp.last_hidden_state = outp.last_hidden_state.detach().requires_grad_(True)
        outp.pooler_output = outp.pooler_output.detach().requires_grad_(True)
        cap["text"].append(outp)
        return outp

    h1 = model.dual_encoder.text_model.register_forward_hook(text_hook)
    # vision tower is frozen (mm_late.py:67-69): its outputs carry no gradient
    for prm in model.dual_encoder.vision_model.parameters():
        assert not prm.requires_grad
    h2 = model.dual_encoder.vision_model.register_forward_hook(lambda _m, _i, o: cap["vision"].append(o))

    tim_inputs, lbl_tim, src = None, None, None
    if use_itm:
        np.random.seed(seed)
        tim_ids, tim_mask, lbl_tim = ref_shims.ref_prepare_itm_inputs(ns, ids, mask)
        np.random.seed(seed)
        swap, src = R.itm_decisions_from_stream(B, np.random)
        assert np.array_equal(tim_ids.numpy(), ids.numpy()[src])
        tim_inputs = (tim_ids, tim_mask)

    out_cls, logits_per_text, out_tim, out_iadds, mm_features = model(ids, mask, pixels, tim_inputs=tim_inputs,
                                                                       iadds_task=False)
    loss_fn = torch.nn.CrossEntropyLoss(weight=class_w)          # run_mm_late.py:85
    tim_loss_fn = torch.nn.CrossEntropyLoss()                    # run_mm_late.py:97
    beta_itc, beta_itm = 0.1, 0.1
    label = y_soft.type_as(out_cls)                              # mm_late.py:471
    if use_itm:
        loss = (1 - (beta_itc + beta_itm)) * loss_fn(out_cls, label) + beta_itc * ns.utils.clip_loss(logits_per_text) \
            + beta_itm * tim_loss_fn(out_tim, lbl_tim)           # mm_late.py:474
    else:
        loss = (1 - beta_itc) * loss_fn(out_cls, label) + beta_itc * ns.utils.clip_loss(logits_per_text)  # :476
    loss.backward()
    h1.remove()
    h2.remove()

    t1, v1 = cap["text"][0], cap["vision"][0]
    out = {
        "fusion": np.array(fusion), "use_itm": np.array(use_itm), "seed": np.array(seed), "C": np.array(C),
        "beta_itc": np.array(beta_itc), "beta_itm": np.array(beta_itm),
        "x_t": t1.last_hidden_state.detach().numpy(), "t_pool": t1.pooler_output.detach().numpy(),
        "x_v": v1.last_hidden_state.detach().numpy(), "v_pool": v1.pooler_output.detach().numpy(),
        "y_soft": y_soft.numpy(), "class_w": class_w.numpy(),
        "out_cls": out_cls.detach().numpy(), "logits_per_text": logits_per_text.detach().numpy(),
        "mm_features": mm_features.detach().numpy(), "loss": loss.detach().numpy(),
        "loss_cls": loss_fn(out_cls, label).detach().numpy(),
        "loss_itc": ns.utils.clip_loss(logits_per_text).detach().numpy(),
    }
    zeros = torch.zeros_like(t1.last_hidden_state)
    out["d_x_t_pass1"] = (t1.last_hidden_state.grad if t1.last_hidden_state.grad is not None else zeros).numpy()
    out["d_t_pool_pass1"] = (t1.pooler_output.grad if t1.pooler_output.grad is not None
                             else torch.zeros_like(t1.pooler_output)).numpy()
    if use_itm:
        t2 = cap["text"][1]
        out["lbl_tim"] = lbl_tim.numpy()
        out["src_idx"] = np.asarray(src)
        out["out_tim"] = out_tim.detach().numpy()
        out["loss_itm"] = tim_loss_fn(out_tim, lbl_tim).detach().numpy()
        out["x_t_tim"] = t2.last_hidden_state.detach().numpy()
        out["d_x_t_pass2"] = (t2.last_hidden_state.grad if t2.last_hidden_state.grad is not None else zeros).numpy()
    named = dict(model.named_parameters())
    for k in HEAD_PARAM_NAMES:
        g = named[k].grad
        if g is None:
            out["gnone_" + k] = np.array(1)
            continue
        if g.numel() <= 4096:
            out["g_" + k] = g.numpy()
        else:
            out["gsum_" + k], out["gblk_" + k] = grad_summary(g)
    name = "head_%s%s.npz" % (fusion.replace("-", ""), "_itm" if use_itm else "")
    np.savez_compressed(os.path.join(GOLD, name), **out)
    print(name, "loss", float(loss), "size KB", os.path.getsize(os.path.join(GOLD, name)) // 1024)


def gen_mm_early(ns):
    out = {}
    for seed, B, d in ((0, 8, 768), (1, 33, 64), (2, 48, 256)):
        g = torch.Generator().manual_seed(seed)
        T = torch.randn(B, d, generator=g).requires_grad_(True)
        V = (torch.randn(B, d, generator=g) + 0.3 * T.detach()).requires_grad_(True)
        W = torch.randn(B, B, generator=g)
        ls = torch.tensor(2.6592, requires_grad=True)
        S = ref_shims.ref_early_logits(ns, T, V, ls)                       # mm_early.py:96-103
        closs = ns.utils.clip_loss(S)                                      # utils.py:228-231 (mm_early.py:367-368)
        (closs + 1e-3 * (S * W).sum()).backward()
        key = "b%d_d%d" % (B, d)
        for k, v in (("T", T), ("V", V), ("W", W), ("S", S), ("clip_loss", closs), ("dT", T.grad), ("dV", V.grad),
                     ("dls", ls.grad)):
            out[key + "_" + k] = v.detach().numpy()
    for seed, B in ((40, 8), (30, 16), (7, 1)):
        L = 6
        ids = (torch.arange(B * L).view(B, L) * 7 + 3) % 1000
        mask = ((torch.arange(B * L).view(B, L) % 5) != 0).long()
        tt = (torch.arange(B * L).view(B, L) % 2)
        np.random.seed(seed)
        tim_ids, tim_mask, tim_tt, lbl = ref_shims.ref_early_prepare_itm_inputs(ns, ids, mask, tt)   # mm_early.py:262-293
        key = "s%d_b%d" % (seed, B)
        for k, v in (("ids", ids), ("mask", mask), ("tt", tt), ("tim_ids", tim_ids), ("tim_mask", tim_mask), ("tim_tt", tim_tt),
                     ("lbl", lbl)):
            out[key + "_" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, "mm_early_tail.npz"), **out)
    print("mm_early_tail: clip_loss", float(out["b8_d768_clip_loss"]), "lbl", out["s40_b8_lbl"].tolist())


def main():
    os.makedirs(GOLD, exist_ok=True)
    ns = ref_shims.load_reference(small_encoders=True)
    gen_mm_early(ref_shims.load_reference_early())
    if "--only-early" in sys.argv:
        return
    gen_itm_stream(ns)
    gen_clip_loss(ns)
    for fusion, itm in (("concat", True), ("attention", True), ("gmu", True), ("aspect-att", False), ("concat", False)):
        gen_head(ns, fusion, itm)


if __name__ == "__main__":
    main()
