"""GPU parity tests added in round 2 (run with -m gpu on the B200 box), all through the C ABI:

  * hard-negative ITM sampling in TILE-STREAM form (tic_itc_fwd(qpart) -> tic_itm_hard_locate -> tic_itc_pick): the
    oracle's itm_sample_hard fed the CUDA-produced logits gives identical labels / source rows (bit-exact bar), alone and
    end to end inside a HeadPlan step (losses / gradients <= 1e-3);
  * the captured step as a TRAINING step: one CUDA graph replayed across optimiser steps follows the fp32 master weights
    and the trainable logit_scale (device scalar), matching the oracle's trajectory;
  * the sizes the roofline numbers are quoted at: ITC at 16384 x 768 and 65536 x 256 against a chunked torch-fp32
    evaluation on the GPU, attention fusion at B = 4096 in the default (split-precision) mode against the fp64 oracle;
  * the forward/backward generation check of the autograd mode.
"""
import math

import numpy as np
import pytest
import torch

from oracle import restatement as R
from test_gpu_parity import REL, _bf, _check_head, _dev, _make_head_case, _plan_mod, _rel

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------------- hard negatives
@pytest.mark.parametrize("B,Pd,split", [(2, 64, False), (64, 64, True), (256, 512, True), (1000, 768, False), (2048, 256, False),
                                        (2500, 512, False), (4096, 512, False)])
def test_hard_tile_stream_bit_exact(B, Pd, split):
    """narrow tiles (B <= 2048), wide tiles with a ragged edge (2500), 2-CTA multicast tiles (4096), split-precision
    operands (extra K-segments).  Oracle input = the logits the SAME forward launch materialised."""
    P = _plan_mod()
    g = torch.Generator().manual_seed(B + Pd)
    T32 = torch.randn(B, Pd, generator=g)
    V32 = torch.randn(B, Pd, generator=g) + 0.5 * T32
    Td, Vd = T32.to(torch.bfloat16).to(_dev()), V32.to(torch.bfloat16).to(_dev())
    Tl = (T32 - Td.float().cpu()).to(torch.bfloat16).to(_dev()) if split else None
    Vl = (V32 - Vd.float().cpu()).to(torch.bfloat16).to(_dev()) if split else None
    rs = np.random.RandomState(B)
    u_coin, u_pick = rs.uniform(size=B).astype(np.float32), rs.uniform(size=B).astype(np.float32)
    uc, up = torch.tensor(u_coin, device=_dev()), torch.tensor(u_pick, device=_dev())
    scale = float(np.float32(math.exp(2.6592)))
    results = []
    for materialize in (True, False):
        it = P.ItcPlan(B, B, Pd, _dev(), materialize_logits=materialize, hard=True)
        lbl = torch.full((B,), -7, dtype=torch.int64, device=_dev())
        src = torch.full((B,), -7, dtype=torch.int32, device=_dev())
        it.norms(Td, Pd, Vd, Pd, T_lo=Tl, V_lo=Vl)
        it.fwd_tiles(Td, Pd, Vd, Pd, scale, T_lo=Tl, V_lo=Vl)
        it.hard_locate(uc, up, lbl, src)
        it.hard_pick(Td, Pd, Vd, Pd, scale, src, T_lo=Tl, V_lo=Vl)
        torch.cuda.synchronize()
        results.append((lbl.cpu().numpy(), src.cpu().numpy().astype(np.int64)))
        if materialize:
            S = it.logits.cpu().numpy()
            lbl_o, src_o = R.itm_sample_hard(S, u_coin, u_pick, ref=np.float32(scale))
            assert np.array_equal(results[0][0], lbl_o)
            assert np.array_equal(results[0][1], src_o), "mismatching rows: %s" % np.nonzero(results[0][1] != src_o)[0][:10]
            # the materialised-S kernel implements the same spec
            lbl2 = torch.empty(B, dtype=torch.int64, device=_dev())
            src2 = torch.empty(B, dtype=torch.int32, device=_dev())
            from tic_b200 import capi
            capi.call("tic_itm_sample", uc.data_ptr(), up.data_ptr(), B, 1, it.logits.data_ptr(), B, scale, None,
                      lbl2.data_ptr(), src2.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert np.array_equal(src2.cpu().numpy().astype(np.int64), src_o)
            if B > 2:   # the sampler is similarity-weighted: mismatch rows differ from the uniform rule's picks somewhere
                _, src_u = R.itm_sample_uniform(u_coin, u_pick)
                assert (src_u != src_o).any()
    # without materialised logits (what the fused step runs) the picks are the same
    assert np.array_equal(results[0][0], results[1][0]) and np.array_equal(results[0][1], results[1][1])


@pytest.mark.parametrize("fusion,B,Lv", [("concat", 256, 2), ("attention", 512, 197)])
def test_head_step_hard_negatives_e2e(fusion, B, Lv):
    """HeadPlan(itm_mode="hard") end to end: source rows = oracle sampler on the CUDA logits (bit-exact), then the whole
    step (losses, outputs, every gradient) against the fp64 oracle run with those decisions."""
    P = _plan_mod()
    C = 4
    dev_in, ora_in = _make_head_case(B, C, 2, Lv, seed=B + 3)
    p32 = R.init_params(C, seed=11)
    plan = P.HeadPlan(B, C=C, fusion=fusion, use_itc=True, use_itm=True, Lv=Lv, itm_mode="hard", materialize_logits=True)
    plan.set_weights(p32)
    out = plan.step(dev_in)
    torch.cuda.synchronize()
    S = out["logits_per_text"].cpu().numpy()
    lbl_o, src_o = R.itm_sample_hard(S, dev_in["u_coin"].cpu().numpy(), dev_in["u_pick"].cpu().numpy(),
                                     ref=np.float32(plan.scale))
    assert np.array_equal(out["lbl_tim"].cpu().numpy(), lbl_o)
    assert np.array_equal(out["src_idx"].cpu().numpy().astype(np.int64), src_o)
    assert (src_o != R.itm_sample_uniform(dev_in["u_coin"].cpu().numpy(), dev_in["u_pick"].cpu().numpy())[1]).any()
    ora_in["lbl_tim"], ora_in["src_idx"] = torch.from_numpy(lbl_o), torch.from_numpy(src_o)
    _check_head(plan, dev_in, ora_in, p32, fusion, True, odev="cuda")
    # without materialised logits (bench configuration) and with ids / mask to gather: same picks, gathered rows exact
    plan2 = P.HeadPlan(B, C=C, fusion=fusion, use_itc=True, use_itm=True, Lv=Lv, itm_mode="hard")
    plan2.set_weights(p32)
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, 30000, (B, 128), generator=g).to(_dev())
    mask = (torch.rand(B, 128, generator=g) > 0.3).long().to(_dev())
    out2 = plan2.step(dict(dev_in, ids=ids, mask=mask))
    torch.cuda.synchronize()
    assert np.array_equal(out2["src_idx"].cpu().numpy().astype(np.int64), src_o)
    assert torch.equal(out2["tim_ids"].cpu(), ids.cpu()[torch.from_numpy(src_o)])
    assert torch.equal(out2["tim_mask"].cpu(), mask.cpu()[torch.from_numpy(src_o)])
    assert abs(float(out2["loss"][0]) - float(out["loss"][0])) <= 1e-6 * abs(float(out["loss"][0]))


# ---------------------------------------------------------------------------------------------------- training step
def _ste_bf16(w):
    """bf16 rounding with a straight-through gradient: what the tensor cores see of an fp32 master weight"""
    return w + (w.detach().float().to(torch.bfloat16).to(w.dtype) - w.detach())


def test_captured_step_is_a_training_step():
    """ONE CUDA graph of HeadPlan.step replayed across 3 optimiser steps: the graph reads the fp32 master weights and the
    trainable logit_scale (mm_late.py:59-69) through pointers (tic_refresh_weights is its first node), so every replay
    uses the updated parameters.  Oracle: the same 3 AdamW steps in fp64 with bf16-rounded (straight-through) matrices."""
    P = _plan_mod()
    B, C, fusion = 64, 4, "concat"
    dev_in, ora_in = _make_head_case(B, C, 2, 2, seed=91)
    p32 = R.init_params(C, seed=5)
    names = ["dual_encoder.text_projection.weight", "dual_encoder.visual_projection.weight", "dual_encoder.logit_scale",
             "linear_fusion.weight", "linear_fusion.bias", "linear_cls.weight", "linear_cls.bias", "linear_tim.weight",
             "linear_tim.bias"]
    grads = {"dual_encoder.text_projection.weight": "dW_t", "dual_encoder.visual_projection.weight": "dW_v",
             "dual_encoder.logit_scale": "d_logit_scale", "linear_fusion.weight": "dW_f", "linear_fusion.bias": "db_f",
             "linear_cls.weight": "dW_cls", "linear_cls.bias": "db_cls", "linear_tim.weight": "dW_tim", "linear_tim.bias": "db_tim"}
    dparams = {k: torch.nn.Parameter(v.clone().to(_dev())) for k, v in p32.items()}
    oparams = {k: v.clone().double().requires_grad_(True) for k, v in p32.items()}
    lr_w, lr_s = 2e-3, 5e-2

    def groups(pp):
        return [{"params": [pp[n] for n in names if n != "dual_encoder.logit_scale"], "lr": lr_w},
                {"params": [pp["dual_encoder.logit_scale"]], "lr": lr_s}]

    dopt = torch.optim.AdamW(groups(dparams), weight_decay=0.01)
    oopt = torch.optim.AdamW(groups(oparams), weight_decay=0.01)
    plan = P.HeadPlan(B, C=C, fusion=fusion, use_itc=True, use_itm=True, Lv=2)
    plan.bind_params(dparams, live=True)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        plan.step(dev_in)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        plan.step(dev_in)
    big = {"dual_encoder.text_projection.weight", "dual_encoder.visual_projection.weight", "linear_fusion.weight"}
    losses_d, losses_o, scales_d, scales_o = [], [], [], []
    for it in range(3):
        graph.replay()
        torch.cuda.synchronize()
        losses_d.append(float(plan.out["loss"][0]))
        scales_d.append(float(dparams["dual_encoder.logit_scale"]))
        for n in names:
            g = plan.out[grads[n]]
            dparams[n].grad = g.reshape(dparams[n].shape).clone()
        dopt.step()
        # oracle
        used = {k: (_ste_bf16(v) if k in big else v) for k, v in oparams.items()}
        ref = R.head_step(dict(ora_in), used, fusion_name=fusion, use_itc=True, use_itm=True, beta_itc=0.1, beta_itm=0.1)
        oopt.zero_grad()
        ref["loss"].backward()
        losses_o.append(float(ref["loss"]))
        scales_o.append(float(oparams["dual_encoder.logit_scale"]))
        oopt.step()
    for a, b in zip(losses_d, losses_o):
        assert abs(a - b) <= REL * abs(b), (losses_d, losses_o)
    for a, b in zip(scales_d, scales_o):
        assert abs(a - b) <= REL * abs(b), (scales_d, scales_o)
    assert abs(float(dparams["dual_encoder.logit_scale"]) - float(oparams["dual_encoder.logit_scale"])) < 1e-3
    # the test is sensitive: the loss moved by far more than the tolerance, and logit_scale moved by ~3 * lr_s
    assert abs(losses_d[2] - losses_d[0]) > 20 * REL * abs(losses_d[0])
    assert abs(scales_d[2] - scales_d[0]) > 0.02          # ~2 Adam steps of lr_s = 0.05 (a stale graph would keep the scale)
    assert int(plan.scale_status.item()) == 0


def test_autograd_generation_check():
    """A backward that follows ANOTHER forward on the same plan must fail loudly (the activations live in the plan)."""
    P = _plan_mod()
    B, C = 8, 4
    dev_in, ora_in = _make_head_case(B, C, 2, 2, seed=3)
    plan = P.HeadPlan(B, C=C, fusion="concat", use_itc=True, use_itm=False, Lv=2, materialize_logits=True)
    plan.set_weights(R.init_params(C, seed=1))
    plan.forward(dev_in)
    g1 = plan.generation
    d_cls = torch.ones(B, C, device=_dev())
    plan.backward(dev_in, d_out_cls=d_cls, generation=g1)      # matching forward: fine
    plan.forward(dev_in)
    plan.forward(dev_in)
    with pytest.raises(RuntimeError):
        plan.backward(dev_in, d_out_cls=d_cls, generation=g1)


# ---------------------------------------------------------------------------------------------------- BASELINE sizes
def _torch_fp32_itc_chunked(T, V, s, chunk=4096):
    """clip_loss(s * That Vhat^T) and its gradients w.r.t. T, V, log-scale in plain torch fp32 (TF32 off) on the GPU,
    row-chunked so that nothing of size B^2 is held: two passes (statistics, then gradients)."""
    B = T.shape[0]
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        Tf, Vf = T.float(), V.float()
        nt, nv = Tf.norm(dim=1, keepdim=True), Vf.norm(dim=1, keepdim=True)
        Tn, Vn = Tf / nt, Vf / nv
        lse_r = torch.empty(B, device=T.device)
        col_m = torch.full((B,), -float("inf"), device=T.device)
        col_s = torch.zeros(B, device=T.device)
        diag = (s * (Tn * Vn).sum(1))
        for i in range(0, B, chunk):
            S = s * Tn[i:i + chunk] @ Vn.t()
            lse_r[i:i + chunk] = torch.logsumexp(S, 1)
            m = torch.maximum(col_m, S.max(0).values)
            col_s = col_s * torch.exp(col_m - m) + torch.exp(S - m).sum(0)
            col_m = m
        lse_c = col_m + torch.log(col_s)
        loss = 0.5 * ((lse_r - diag).mean() + (lse_c - diag).mean())
        dTn, dVn = torch.zeros_like(Tn), torch.zeros_like(Vn)
        dls = torch.zeros((), device=T.device, dtype=torch.float64)
        for i in range(0, B, chunk):
            S = s * Tn[i:i + chunk] @ Vn.t()
            G = (torch.exp(S - lse_r[i:i + chunk, None]) + torch.exp(S - lse_c[None, :])) / (2 * B)
            idx = torch.arange(i, min(i + chunk, B), device=T.device)
            G[idx - i, idx] -= 1.0 / B
            dTn[i:i + chunk] = s * (G @ Vn)
            dVn += s * (G.t() @ Tn[i:i + chunk])
            dls += (G.double() * S.double()).sum()
        dT = (dTn - Tn * (Tn * dTn).sum(1, keepdim=True)) / nt
        dV = (dVn - Vn * (Vn * dVn).sum(1, keepdim=True)) / nv
        return float(loss), dT, dV, float(dls)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("B,Pd", [(16384, 768), (65536, 256)])
def test_itc_roofline_sizes_vs_torch_fp32(B, Pd):
    """The sizes every roofline number is quoted at (65536: row * ld_ga reaches 2^32 — the index-width spot check)."""
    P = _plan_mod()
    g = torch.Generator().manual_seed(B + 1)
    T = torch.randn(B, Pd, generator=g).to(torch.bfloat16).to(_dev())
    V = (torch.randn(B, Pd, generator=g) + 0.4 * T.cpu().float()).to(torch.bfloat16).to(_dev())
    s = float(np.exp(2.6592))
    plan = P.ItcPlan(B, B, Pd, _dev())
    sums, rsum = torch.zeros(2, device=_dev()), torch.zeros(1, device=_dev())
    dT, dV = torch.empty(B, Pd, device=_dev()), torch.empty(B, Pd, device=_dev())
    plan.run(T, V, s, 1.0, sums, rsum, dT_f32=dT, dV_f32=dV)
    torch.cuda.synchronize()
    loss = float(0.5 * (sums[0] + sums[1]) / B)
    del plan
    torch.cuda.empty_cache()
    ref, rT, rV, rls = _torch_fp32_itc_chunked(T, V, s)
    assert abs(loss - ref) / abs(ref) < REL, (loss, ref)
    eT, eV = _rel(dT, rT), _rel(dV, rV)
    assert eT < REL and eV < REL, (eT, eV)
    # per-row gradient norms (a max-norm over the whole matrix could hide a bad row block, e.g. an index overflow)
    nT = (dT.norm(dim=1) - rT.norm(dim=1)).abs() / rT.norm(dim=1).clamp_min(1e-12)
    nV = (dV.norm(dim=1) - rV.norm(dim=1)).abs() / rV.norm(dim=1).clamp_min(1e-12)
    assert float(nT.max()) < 5e-3 and float(nV.max()) < 5e-3, (float(nT.max()), float(nV.max()))
    assert float(nT.mean()) < REL and float(nV.mean()) < REL
    assert abs(float(rsum) - rls) < REL * abs(rls) + 1e-6


def test_attention_fusion_c4_size_default_mode():
    """BASELINE configs[3] size: attention fusion, B = 4096, 197 patches, default split-precision mode, ITM on
    (uniform decisions here; the hard-negative form is covered above) against the fp64 oracle evaluated on the GPU."""
    P = _plan_mod()
    B, C, Lv = 4096, 4, 197
    dev_in, ora_in = _make_head_case(B, C, 1, Lv, seed=4096)
    p32 = R.init_params(C, seed=7)
    plan = P.HeadPlan(B, C=C, fusion="attention", use_itc=True, use_itm=True, Lv=Lv)
    assert plan.split
    plan.set_weights(p32)
    _check_head(plan, dev_in, ora_in, p32, "attention", True, odev="cuda")


# ---------------------------------------------------------------------------------------------------- c2 restructuring
@pytest.mark.parametrize("B", [8, 200, 256])
def test_c2_fused_forms_match_unfused(B, monkeypatch):
    """The launch-count work on the small-batch step: (1) forward + gradient-operand ITC tiles in ONE cluster launch,
    (2) concat fusion in pairwise form (Pt[src] + Pv, no packed operand), (3) weight gradients by plain stores.  Each
    against the round-1 forms (switches TIC_ITC_FUSED_SMALL / TIC_CONCAT_PAIRWISE) on the same inputs, and the new default
    against the fp64 oracle."""
    P = _plan_mod()
    C = 4
    dev_in, ora_in = _make_head_case(B, C, 2, 2, seed=B + 77)
    p32 = R.init_params(C, seed=13)
    new = P.HeadPlan(B, C=C, fusion="concat", use_itc=True, use_itm=True, Lv=2)
    assert new.pairwise and new.fuse_itc_small and new.itc.can_fuse_small
    assert not any(new._atomic.values()), new._atomic          # K = B is short: no split-K, no atomics, no large memset
    new.set_weights(p32)
    _check_head(new, dev_in, ora_in, p32, "concat", True)
    assert new._itc_bwd_fused
    monkeypatch.setenv("TIC_CONCAT_PAIRWISE", "0")
    monkeypatch.setenv("TIC_ITC_FUSED_SMALL", "0")
    old = P.HeadPlan(B, C=C, fusion="concat", use_itc=True, use_itm=True, Lv=2)
    assert not old.pairwise and not old.fuse_itc_small
    old.set_weights(p32)
    a, b = new.step(dev_in), old.step(dev_in)
    torch.cuda.synchronize()
    assert not old._itc_bwd_fused
    for k in ("loss", "out_cls", "out_tim", "mm_features", "dW_f", "db_f", "dW_t", "dW_v", "d_t_pool", "d_xt_cls", "dW_cls", "dW_tim"):
        assert _rel(a[k], b[k]) < 2e-4, k
    assert torch.equal(a["src_idx"], b["src_idx"])
    # the ITC half is the same arithmetic in both forms (same tiles, same epilogues): bit-identical statistics
    assert torch.equal(new.itc.row_part, old.itc.row_part) and torch.equal(new.itc.GA, old.itc.GA)
