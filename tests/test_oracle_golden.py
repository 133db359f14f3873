"""CPU: the oracle restatement against the golden fixtures recorded from the unmodified reference
(oracle/make_golden.py).  These pin the oracle; the -m gpu tests then compare the CUDA path with the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import restatement as R

HEAD_CASES = ["head_concat_itm", "head_attention_itm", "head_gmu_itm", "head_aspectatt", "head_concat"]


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False))


@pytest.mark.parametrize("seed,B", [(40, 8), (30, 16), (123, 256), (0, 2), (7, 1)])
def test_itm_stream_matches_reference(golden_dir, seed, B):
    g = _load(golden_dir, "itm_stream")
    key = "s%d_b%d" % (seed, B)
    ids, mask = torch.from_numpy(g[key + "_ids"]), torch.from_numpy(g[key + "_mask"])
    tim_ids, tim_mask, lbl = R.prepare_itm_inputs_stream(ids, mask, np.random.RandomState(seed))
    assert np.array_equal(tim_ids.numpy(), g[key + "_tim_ids"])
    assert np.array_equal(tim_mask.numpy(), g[key + "_tim_mask"])
    assert np.array_equal(lbl.numpy(), g[key + "_lbl"])
    # the uniform-driven statement (what the CUDA sampler implements) replays the same decisions
    swap, src = R.itm_decisions_from_stream(B, np.random.RandomState(seed))
    assert np.array_equal(src, g[key + "_src"])
    u_coin, u_pick = R.uniforms_from_decisions(swap, src)
    lbl_u, src_u = R.itm_sample_uniform(u_coin, u_pick)
    assert np.array_equal(lbl_u, g[key + "_lbl"])
    assert np.array_equal(src_u, g[key + "_src"])


def test_survey_golden_vectors():
    # np.random.seed(40), B=8 and np.random.seed(30), B=16
    swap, src = R.itm_decisions_from_stream(8, np.random.RandomState(40))
    assert (~swap).astype(int).tolist() == [0, 1, 1, 0, 0, 0, 1, 1]
    assert src.tolist() == [4, 1, 2, 0, 1, 2, 6, 7]
    swap, src = R.itm_decisions_from_stream(16, np.random.RandomState(30))
    assert (~swap).astype(int).tolist() == [1, 1, 1, 1, 0, 1, 0, 1, 1, 0, 1, 1, 1, 1, 1, 1]
    assert src.tolist() == [0, 1, 2, 3, 13, 5, 5, 7, 8, 3, 10, 11, 12, 13, 14, 15]


@pytest.mark.parametrize("B", [1, 2, 8, 33, 128])
def test_clip_loss_matches_reference(golden_dir, B):
    g = _load(golden_dir, "clip_loss")
    S = torch.tensor(g["b%d_S" % B], requires_grad=True)
    loss = R.clip_loss(S)
    loss.backward()
    np.testing.assert_allclose(loss.detach().numpy(), g["b%d_loss" % B], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(S.grad.numpy(), g["b%d_dS" % B], rtol=1e-5, atol=1e-8)


def test_clip_loss_known_answers(golden_dir):
    g = _load(golden_dir, "clip_loss")
    # clip_loss(c*I) = log(1 + (B-1) e^-c); clip_loss(0) = log B
    np.testing.assert_allclose(g["eye8_c5"], np.log1p(7 * np.exp(-5.0)), rtol=1e-6)
    np.testing.assert_allclose(g["eye16_c0"], np.log(16.0), rtol=1e-6)
    np.testing.assert_allclose(R.clip_loss(5.0 * torch.eye(8)).numpy(), g["eye8_c5"], rtol=1e-6)


def _head_inputs(g, dtype=torch.float32, requires_grad=True):
    inp = {k: torch.tensor(g[k], dtype=dtype) for k in ("x_t", "x_v", "t_pool", "v_pool", "y_soft", "class_w")}
    if bool(g["use_itm"]):
        inp["lbl_tim"] = torch.from_numpy(g["lbl_tim"])
        inp["src_idx"] = torch.from_numpy(g["src_idx"])
    if requires_grad:
        inp["x_t"].requires_grad_(True)
        inp["t_pool"].requires_grad_(True)
    return inp


@pytest.mark.parametrize("case", HEAD_CASES)
def test_head_step_matches_reference(golden_dir, case):
    g = _load(golden_dir, case)
    fusion, use_itm, C, seed = str(g["fusion"]), bool(g["use_itm"]), int(g["C"]), int(g["seed"])
    p = R.init_params(C, seed=seed)
    for v in p.values():
        v.requires_grad_(True)
    inp = _head_inputs(g)
    out = R.head_step(inp, p, fusion_name=fusion, use_itc=True, use_itm=use_itm, beta_itc=float(g["beta_itc"]),
                      beta_itm=float(g["beta_itm"]))
    tol = dict(rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(out["logits_per_text"].detach().numpy(), g["logits_per_text"], **tol)
    np.testing.assert_allclose(out["out_cls"].detach().numpy(), g["out_cls"], **tol)
    np.testing.assert_allclose(out["mm_features"].detach().numpy(), g["mm_features"], **tol)
    np.testing.assert_allclose(out["loss_cls"].item(), g["loss_cls"], rtol=1e-5)
    np.testing.assert_allclose(out["loss_itc"].item(), g["loss_itc"], rtol=1e-5)
    np.testing.assert_allclose(out["loss"].item(), g["loss"], rtol=1e-5)
    if use_itm:
        # text_model(ids[src]) == text_model(ids)[src]: the second encoder pass is a row gather
        np.testing.assert_allclose(g["x_t_tim"], g["x_t"][g["src_idx"]], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(out["out_tim"].detach().numpy(), g["out_tim"], **tol)
        np.testing.assert_allclose(out["loss_itm"].item(), g["loss_itm"], rtol=1e-5)
    out["loss"].backward()
    gt = dict(rtol=2e-3, atol=2e-6)
    for k, v in p.items():
        if "g_" + k in g:
            np.testing.assert_allclose(v.grad.numpy(), g["g_" + k], err_msg=k, **gt)
        elif "gsum_" + k in g:
            gs = v.grad.double()
            np.testing.assert_allclose([gs.sum().item(), gs.abs().sum().item(), gs.norm().item()], g["gsum_" + k],
                                       rtol=2e-3, atol=1e-6, err_msg=k)
            np.testing.assert_allclose(gs.reshape(gs.shape[0], -1)[:4, :8].numpy(), g["gblk_" + k], err_msg=k, **gt)
        elif "gnone_" + k in g:
            assert v.grad is None or float(v.grad.abs().sum()) == 0.0, k
    d_xt = torch.from_numpy(g["d_x_t_pass1"]).clone()
    if use_itm:
        d_xt.index_add_(0, torch.from_numpy(g["src_idx"]), torch.from_numpy(g["d_x_t_pass2"]))
    got_dxt = inp["x_t"].grad if inp["x_t"].grad is not None else torch.zeros_like(inp["x_t"])  # aspect-att reads pools only
    np.testing.assert_allclose(got_dxt.numpy(), d_xt.numpy(), **gt)
    np.testing.assert_allclose(inp["t_pool"].grad.numpy(), g["d_t_pool_pass1"], **gt)


def test_attention_collapse_equals_literal(golden_dir):
    g = _load(golden_dir, "head_attention_itm")
    p = R.init_params(int(g["C"]), seed=int(g["seed"]), dtype=torch.float64)
    x_t, x_v = torch.tensor(g["x_t"], dtype=torch.float64), torch.tensor(g["x_v"], dtype=torch.float64)
    a = R.fusion_attention_literal(x_t, x_v, p)
    b = R.fusion_attention_collapsed(x_t[:, 0, :], x_v, p)
    assert float((a - b).abs().max()) < 1e-12


def test_aspect_scramble_is_not_pairing():
    # stack->reshape pairs flat rows (2i, 2i+1), equal to the (t_i, v_i) pairing only for B == 1
    p = R.init_params(4, seed=1)
    rs = np.random.RandomState(0)
    t, v = torch.tensor(rs.normal(size=(3, 768)), dtype=torch.float32), torch.tensor(rs.normal(size=(3, 768)), dtype=torch.float32)
    out = R.fusion_aspect(t, v, p)
    flat = torch.cat((t, v), 0)
    for i in range(3):
        pair = flat[2 * i:2 * i + 2]
        e = torch.tanh(pair @ p["aspectattention.weight"].t() + p["aspectattention.bias"])
        w = torch.softmax(e, dim=0)
        np.testing.assert_allclose(out[i].numpy(), torch.relu((w * pair).sum(0)).numpy(), rtol=1e-5, atol=1e-6)
    with pytest.raises(TypeError):
        R.mm_fusion("aspect-att", t[:, None, :], v[:, None, :], p)  # the reference's ITM branch passes no pools


def test_det_exp_accuracy_and_hard_sampler_properties():
    x = -np.abs(np.random.RandomState(3).normal(0, 8, size=20000)).astype(np.float32)
    y = R.det_exp_f32(x)
    ref = np.exp(x.astype(np.float64))
    assert np.max(np.abs(y - ref) / ref) < 3e-7
    assert R.det_exp_f32(np.zeros(1, np.float32))[0] == np.float32(1.0)
    rs = np.random.RandomState(5)
    B = 64
    S = rs.normal(0, 2, size=(B, B)).astype(np.float32)
    u_coin, u_pick = rs.uniform(size=B).astype(np.float32), rs.uniform(size=B).astype(np.float32)
    lbl, src = R.itm_sample_hard(S, u_coin, u_pick)
    assert np.array_equal(lbl, (u_coin >= 0.5).astype(np.int64))
    assert np.all(src[lbl == 1] == np.arange(B)[lbl == 1])
    assert np.all(src[lbl == 0] != np.arange(B)[lbl == 0])
    # a dominant hard negative is picked for every uniform
    S2 = np.full((4, 4), -30.0, np.float32)
    S2[np.arange(4), (np.arange(4) + 1) % 4] = 10.0
    lbl, src = R.itm_sample_hard(S2, np.zeros(4, np.float32), np.array([0.0, 0.3, 0.6, 0.999], np.float32))
    assert src.tolist() == [1, 2, 3, 0]


# ---------------------------------------------------------------------------------------------------- mm_early tail (§8 f-3)
@pytest.mark.parametrize("B,d", [(8, 768), (33, 64), (48, 256)])
def test_mm_early_logits_and_gradients_match_reference(golden_dir, B, d):
    """ViLT.get_logits_per_text (mm_early.py:96-103) + utils.clip_loss, recorded from the unmodified reference."""
    g = _load(golden_dir, "mm_early_tail")
    key = "b%d_d%d_" % (B, d)
    T = torch.tensor(g[key + "T"], requires_grad=True)
    V = torch.tensor(g[key + "V"], requires_grad=True)
    ls = torch.tensor(2.6592, requires_grad=True)
    S = R.itc_logits(T, V, ls)
    closs = R.clip_loss(S)
    (closs + 1e-3 * (S * torch.tensor(g[key + "W"])).sum()).backward()
    np.testing.assert_allclose(S.detach().numpy(), g[key + "S"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(closs.detach().numpy(), g[key + "clip_loss"], rtol=1e-5)
    np.testing.assert_allclose(T.grad.numpy(), g[key + "dT"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(V.grad.numpy(), g[key + "dV"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(ls.grad.numpy(), g[key + "dls"], rtol=1e-4)


@pytest.mark.parametrize("seed,B", [(40, 8), (30, 16), (7, 1)])
def test_mm_early_itm_stream_matches_reference(golden_dir, seed, B):
    """MMEarly_Model.prepare_itm_inputs (mm_early.py:262-293): the late-fusion rule plus token_type_ids, same numpy stream."""
    g = _load(golden_dir, "mm_early_tail")
    key = "s%d_b%d_" % (seed, B)
    swap, src = R.itm_decisions_from_stream(B, np.random.RandomState(seed))
    assert np.array_equal((~swap).astype(np.int64), g[key + "lbl"])
    for name in ("ids", "mask", "tt"):
        assert np.array_equal(R.gather_rows(torch.from_numpy(g[key + name]), src).numpy(), g[key + "tim_" + name])
    late = _load(golden_dir, "itm_stream")      # the two wrappers consume the stream identically
    assert np.array_equal(g[key + "lbl"], late["s%d_b%d_lbl" % (seed, B)])
