"""GPU tests (run with -m gpu) of the C-ABI entry points added for the small-batch fusions, the peer-memory exchange, the
pipelined host entry point and the large-batch kernels — each against the CPU oracle (oracle/restatement.py) or against
the unfused entry point it replaces, through ctypes -> libtic_b200.so.

Tolerances: integer / index work bit-exact; fp32 statistics 1e-5 relative; ITC losses and gradients 1e-3 (north_star)."""
import ctypes
import math

import numpy as np
import pytest
import torch

from oracle import restatement as R

pytestmark = pytest.mark.gpu
REL = 1e-3


def _dev():
    return torch.device("cuda:0")


def _rel(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def _mods():
    import tic_b200.plan as P
    from tic_b200 import capi
    return P, capi


# ---------------------------------------------------------------------------------------------------- fused L2-norm statistics
@pytest.mark.parametrize("M,N,K", [(256, 512, 768), (100, 200, 136), (64, 64, 64)])
def test_gemm_rowss_matches_row_rnorm(M, N, K):
    """tic_gemm_bf16_rowss: the per-tile row sums of squares reproduce 1/||row|| of the GEMM result (HF :268-269)."""
    P, capi = _mods()
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).to(_dev())
    W = torch.randn(N, K, generator=g).to(torch.bfloat16).to(_dev())
    ldn = (N + 7) // 8 * 8
    D = torch.zeros(M, ldn, dtype=torch.bfloat16, device=_dev())
    D_lo = torch.zeros(M, ldn, dtype=torch.bfloat16, device=_dev())
    nss = capi.load().tic_gemm_rowss_parts(N)
    assert nss == (N + 63) // 64
    ss = torch.full((nss, M), float("nan"), device=_dev())
    P.gemm(A, K, 0, W, K, 0, D, ldn, 1, M, N, K, D_lo=D_lo, row_ss=ss)
    torch.cuda.synchronize()
    ref = A.double().cpu() @ W.double().cpu().t()
    assert _rel(ss.sum(0), (ref ** 2).sum(1)) < 1e-5
    # and against the separate norm kernel on the (hi, lo) pair the GEMM wrote
    rinv = torch.empty(M, device=_dev())
    capi.call("tic_row_rnorm_bf16", D.data_ptr(), D_lo.data_ptr(), ldn, M, N, rinv.data_ptr(), None, 0,
              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert _rel(1.0 / torch.sqrt(ss.sum(0)), rinv) < 1e-4


@pytest.mark.parametrize("B", [64, 256, 1000])
def test_itc_fused_norm_and_inline_lse_match_unfused(B):
    """tic_itc_fwd with sum-of-squares partials (rinv derived and written in-kernel) and tic_itc_bwd_g with inline lse
    give the same loss and gradient operands as the unfused sequence norm -> tiles -> lse -> recompute."""
    P, capi = _mods()
    Pd = 512
    g = torch.Generator().manual_seed(B)
    X = torch.randn(2 * B, 768, generator=g).to(torch.bfloat16).to(_dev())
    W = (0.05 * torch.randn(Pd, 768, generator=g)).to(torch.bfloat16).to(_dev())
    Y, Y_lo = torch.empty(2 * B, Pd, dtype=torch.bfloat16, device=_dev()), torch.empty(2 * B, Pd, dtype=torch.bfloat16, device=_dev())
    ss = torch.empty(capi.load().tic_gemm_rowss_parts(Pd), 2 * B, device=_dev())
    P.gemm(X, 768, 0, W, 768, 0, Y, Pd, 1, 2 * B, Pd, 768, D_lo=Y_lo, row_ss=ss)
    T, V, Tl, Vl = Y[:B], Y[B:], Y_lo[:B], Y_lo[B:]
    scale = math.exp(2.6592)
    res = []
    for fused in (False, True):
        it = P.ItcPlan(B, B, Pd, _dev())
        sums = torch.zeros(2, device=_dev())
        if fused:
            it.fwd_tiles(T, Pd, V, Pd, scale, T_lo=Tl, V_lo=Vl, ss_t=ss[:, :B].contiguous(), ss_v=ss[:, B:].contiguous())
        else:
            it.norms(T, Pd, V, Pd, T_lo=Tl, V_lo=Vl)
            it.fwd_tiles(T, Pd, V, Pd, scale, T_lo=Tl, V_lo=Vl)
        it.lse_loss(scale, sums)
        assert it.can_inline_lse
        it.bwd_operands(T, Pd, V, Pd, scale, 1.0 / (2 * B), T_lo=Tl, V_lo=Vl, inline_lse=fused)
        torch.cuda.synchronize()
        res.append((sums.clone(), it.rinv_t.clone(), it.rinv_v.clone(), it.GA.float().clone(), it.GBT.float().clone()))
    (s0, rt0, rv0, ga0, gb0), (s1, rt1, rv1, ga1, gb1) = res
    assert _rel(rt1, rt0) < 1e-4 and _rel(rv1, rv0) < 1e-4
    assert _rel(s1, s0) < 1e-4
    assert _rel(ga1[:, :B], ga0[:, :B]) < 1e-2 and _rel(gb1[:, :B], gb0[:, :B]) < 1e-2   # bf16 operands: 2^-8 per element


# ---------------------------------------------------------------------------------------------------- symmetric-mode statistics
def test_itc_lse_rows_matches_oracle():
    P, capi = _mods()
    m, nparts = 300, 5
    g = torch.Generator().manual_seed(3)
    pa, pb = torch.rand(nparts, m, generator=g) + 0.1, torch.rand(nparts, m, generator=g) + 0.1
    diag = torch.randn(m, generator=g)
    shift = 14.0
    d = _dev()
    la, lb, sums = torch.empty(m, device=d), torch.empty(m, device=d), torch.zeros(2, device=d)
    ws = torch.zeros(int(capi.load().tic_itc_lse_rows_workspace_bytes(m)) // 4, device=d)
    pad, pbd, dd = pa.to(d), pb.to(d), diag.to(d)     # keep the device copies alive across the asynchronous launches
    for _ in range(2):   # twice: the workspace ticket must reset itself
        sums.zero_()
        capi.call("tic_itc_lse_rows", pad.data_ptr(), pbd.data_ptr(), nparts, m, dd.data_ptr(), shift,
                  la.data_ptr(), lb.data_ptr(), sums.data_ptr(), ws.data_ptr(), None, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ra, rb = shift + torch.log(pa.double().sum(0)), shift + torch.log(pb.double().sum(0))
        assert _rel(la, ra) < 1e-6 and _rel(lb, rb) < 1e-6
        assert abs(float(sums[0]) - float((ra - diag.double()).sum())) < 1e-2
        assert abs(float(sums[1]) - float((rb - diag.double()).sum())) < 1e-2


# ---------------------------------------------------------------------------------------------------- pack with the inline ITM rule
@pytest.mark.parametrize("B", [1, 7, 256])
def test_pack_inline_rule_is_bit_exact(B):
    """tic_pack_cls_pairs evaluating the uniform ITM rule in place == packing with src_idx from tic_itm_sample == the
    oracle's rule (mm_late.py:396-409) — rows compared bit for bit."""
    P, capi = _mods()
    E, d = 768, _dev()
    rs = np.random.RandomState(B)
    u_coin, u_pick = rs.uniform(size=B).astype(np.float32), rs.uniform(size=B).astype(np.float32)
    lbl_ref, src_ref = R.itm_sample_uniform(u_coin, u_pick)
    xt = torch.randn(B, 3, E).to(torch.bfloat16).to(d)
    xv = torch.randn(B, 2, E).to(torch.bfloat16).to(d)
    uc, up = torch.tensor(u_coin, device=d), torch.tensor(u_pick, device=d)
    st = torch.cuda.current_stream().cuda_stream
    X1 = torch.zeros(2 * B, 2 * E, dtype=torch.bfloat16, device=d)
    X2 = torch.zeros_like(X1)
    capi.call("tic_pack_cls_pairs", xt.data_ptr(), 3 * E, xv.data_ptr(), 2 * E, B, E, None, X1.data_ptr(), 2 * E, uc.data_ptr(),
              up.data_ptr(), st)
    src = torch.tensor(src_ref.astype(np.int32), device=d)
    capi.call("tic_pack_cls_pairs", xt.data_ptr(), 3 * E, xv.data_ptr(), 2 * E, B, E, src.data_ptr(), X2.data_ptr(), 2 * E, None,
              None, st)
    torch.cuda.synchronize()
    assert torch.equal(X1.view(torch.int16), X2.view(torch.int16))
    want = torch.cat([torch.cat([xt[:, 0], xv[:, 0]], 1), torch.cat([xt[torch.from_numpy(src_ref).to(d), 0], xv[:, 0]], 1)], 0)
    assert torch.equal(X1.view(torch.int16), want.view(torch.int16))


def test_unpack_accumulates_main_and_scattered_rows():
    P, capi = _mods()
    B, E, d = 50, 768, _dev()
    g = torch.Generator().manual_seed(1)
    dX = torch.randn(2 * B, E, generator=g).to(d)
    src = torch.randint(0, B, (B,), generator=g).to(torch.int32).to(d)
    out = torch.zeros(B, E, device=d)
    capi.call("tic_unpack_cls_grad", dX.data_ptr(), E, None, 0, B, E, src.data_ptr(), out.data_ptr(), E,
              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = dX[:B].double().cpu().clone()
    ref.index_add_(0, src.cpu().long(), dX[B:].double().cpu())
    assert _rel(out, ref) < 1e-6


# ---------------------------------------------------------------------------------------------------- peer exchange, world = 1
def test_peer_exchange_single_rank_pulls_its_own_block():
    """tic_peer_alloc / tic_peer_exchange with world == 1: the barrier is trivially satisfied and the pull is a copy of the
    rank's own published ranges; repeated calls advance the epoch.  (2-GPU runs: tests/dist_gpu_check.py --mode peer.)"""
    P, capi = _mods()
    d = _dev()
    nbytes, flag_off = 4096, 4096
    p = ctypes.c_void_p()
    capi.call("tic_peer_alloc", nbytes + 256, ctypes.byref(p))
    try:
        from tic_b200.peer import _RawCuda
        block = torch.as_tensor(_RawCuda(int(p.value), nbytes + 256), device=d)
        assert block.data_ptr() == int(p.value)
        block[:nbytes].copy_(torch.arange(nbytes, dtype=torch.int64).to(torch.uint8))
        hb = capi.load().tic_peer_handle_bytes()
        buf = (ctypes.c_ubyte * hb)()
        capi.call("tic_peer_export", int(p.value), buf)          # exporting a handle works without a peer
        dst1, dst2 = torch.zeros(1024, dtype=torch.uint8, device=d), torch.zeros(512, dtype=torch.uint8, device=d)
        ctr = torch.zeros(2, dtype=torch.int32, device=d)
        bases = (ctypes.c_void_p * 1)(int(p.value))
        so, nb = (ctypes.c_int64 * 2)(256, 2048), (ctypes.c_int64 * 2)(1024, 512)
        dst = (ctypes.c_void_p * 2)(dst1.data_ptr(), dst2.data_ptr())
        ds = (ctypes.c_int64 * 2)(1024, 512)
        for k in range(3):
            capi.call("tic_peer_exchange", bases, 1, 0, flag_off, ctr.data_ptr(), 2, so, nb, dst, ds,
                      torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert int(ctr[0]) == k + 1 and int(ctr[1]) == 0
        assert torch.equal(dst1, block[256:1280]) and torch.equal(dst2, block[2048:2560])
        with pytest.raises(capi.TicError):     # misaligned segment is rejected before any launch
            so_bad = (ctypes.c_int64 * 2)(250, 2048)
            capi.call("tic_peer_exchange", bases, 1, 0, flag_off, ctr.data_ptr(), 2, so_bad, nb, dst, ds, None)
    finally:
        capi.call("tic_peer_free", p)


def test_peer_pull_feeds_segment_waiting_tiles_single_rank():
    """tic_peer_pull on a side stream + tic_itc_fwd(seg_ready=...) on the main stream, world == 1: the tile kernel's producer
    waits for the ready word of the (single) segment, which the pull publishes after copying V out of the peer block; the
    statistics equal those of the plain call on the same V."""
    P, capi = _mods()
    d = _dev()
    B, Pd = 4096, 256
    g = torch.Generator().manual_seed(5)
    T = torch.randn(B, Pd, generator=g).to(torch.bfloat16).to(d)
    V = (torch.randn(B, Pd, generator=g)).to(torch.bfloat16).to(d)
    nbytes = B * Pd * 2
    p = ctypes.c_void_p()
    capi.call("tic_peer_alloc", nbytes + 256, ctypes.byref(p))
    try:
        from tic_b200.peer import _RawCuda
        block = torch.as_tensor(_RawCuda(int(p.value), nbytes + 256), device=d)
        block[:nbytes].view(torch.bfloat16).view(B, Pd).copy_(V)
        bases = (ctypes.c_void_p * 1)(int(p.value))
        ctr = torch.zeros(2, dtype=torch.int32, device=d)
        z = (ctypes.c_int64 * 0)()
        capi.call("tic_peer_exchange", bases, 1, 0, nbytes, ctr.data_ptr(), 0, z, z, (ctypes.c_void_p * 0)(), z,
                  torch.cuda.current_stream().cuda_stream)           # the ordering exchange: epoch 1
        ready, tickets = torch.zeros(1, dtype=torch.int32, device=d), torch.zeros(1, dtype=torch.int32, device=d)
        V_all = torch.zeros(B, Pd, dtype=torch.bfloat16, device=d)
        scale = math.exp(2.6592)
        ref = P.ItcPlan(B, B, Pd, d)
        ref.norms(T, Pd, V, Pd)
        ref.fwd_tiles(T, Pd, V, Pd, scale)
        it = P.ItcPlan(B, B, Pd, d)
        it.norm_t(T, Pd)
        it.rinv_v.copy_(ref.rinv_v)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        so, nb = (ctypes.c_int64 * 1)(0), (ctypes.c_int64 * 1)(nbytes)
        dst, ds = (ctypes.c_void_p * 1)(V_all.data_ptr()), (ctypes.c_int64 * 1)(nbytes)
        with torch.cuda.stream(side):
            capi.call("tic_peer_pull", bases, 1, 0, ctr.data_ptr(), ready.data_ptr(), tickets.data_ptr(), 1, so, nb, dst, ds, 0,
                      side.cuda_stream)
        it.fwd_tiles(T, Pd, V_all, Pd, scale, seg=(ready, ctr, B, 0))
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        assert int(ready[0]) == 1 and torch.equal(V_all.view(torch.int16), V.view(torch.int16))
        assert _rel(it.row_part.sum(0), ref.row_part.sum(0)) < 1e-5
        assert _rel(it.col_part.sum(0), ref.col_part.sum(0)) < 1e-5
        assert _rel(it.diag, ref.diag) < 1e-6
    finally:
        capi.call("tic_peer_free", p)


# ---------------------------------------------------------------------------------------------------- pipelined host entry point
def test_host_pipeline_matches_synchronous_host_step():
    P, capi = _mods()
    B, C, E = 64, 4, 768
    g = torch.Generator().manual_seed(40)
    host = {"x_t": torch.randn(B, 1, E, generator=g), "x_v": torch.randn(B, 1, E, generator=g),
            "t_pool": torch.tanh(torch.randn(B, E, generator=g)), "v_pool": torch.tanh(torch.randn(B, E, generator=g)),
            "y_soft": torch.eye(C)[torch.randint(0, C, (B,), generator=g)], "u_coin": torch.rand(B, generator=g),
            "u_pick": torch.rand(B, generator=g)}
    bfk = ("x_t", "x_v", "t_pool", "v_pool")
    plan = P.HeadPlan(B, C=C, fusion="concat", use_itc=True, use_itm=True, Lv=1, device=_dev())
    plan.set_weights(R.init_params(C, seed=40))
    sync = P.HostStep(plan, host, bf16_keys=bfk)
    want = sync()
    pipe = P.HostPipeline(plan, host, bf16_keys=bfk)
    got = []
    for k in range(5):
        pipe.submit()
        if k >= 1:
            got.append(pipe.result())
    got.append(pipe.result())
    assert len(got) == 5
    for losses in got:
        assert np.allclose(losses, want, rtol=1e-5, atol=1e-6), (losses, want)
    # a different batch in the second slot must produce a different loss, in submission order
    pipe.host_views()["y_soft"].copy_(torch.eye(C)[torch.randint(0, C, (B,), generator=g)])
    pipe.submit()
    pipe.submit()
    a, b = pipe.result(), pipe.result()
    assert not np.allclose(a, b, rtol=1e-6) or np.allclose(a, want, rtol=1e-5)


# ---------------------------------------------------------------------------------------------------- large-batch ITC kernels
@pytest.mark.parametrize("B,Pd", [(4096, 256), (5000, 768)])
def test_itc_large_batch_paths_vs_oracle(B, Pd):
    """From 4096 columns on: 2-CTA multicast tiles, factored one-exponential backward with TMA tile stores of the gradient
    operand, GA-shared image-side product, cluster-multicast GEMMs.  Loss and both gradients against the fp64 oracle."""
    P, capi = _mods()
    g = torch.Generator().manual_seed(B + Pd)
    T32 = torch.randn(B, Pd, generator=g)
    V32 = torch.randn(B, Pd, generator=g) + 0.3 * T32
    Tb, Vb = T32.to(torch.bfloat16), V32.to(torch.bfloat16)
    T, V = Tb.to(_dev()), Vb.to(_dev())
    it = P.ItcPlan(B, B, Pd, _dev())
    assert not it.precise and it.shared_ga
    sums, rsum = torch.zeros(2, device=_dev()), torch.zeros(1, device=_dev())
    dT, dV = torch.empty(B, Pd, device=_dev()), torch.empty(B, Pd, device=_dev())
    scale = math.exp(2.6592)
    it.run(T, V, scale, 1.0, sums, rsum, dT_f32=dT, dV_f32=dV)
    torch.cuda.synchronize()
    Tq, Vq = Tb.double().requires_grad_(True), Vb.double().requires_grad_(True)
    ls = torch.tensor(2.6592, dtype=torch.float64, requires_grad=True)
    ref = R.clip_loss(R.itc_logits(Tq, Vq, ls))
    ref.backward()
    loss = 0.5 * (float(sums[0]) + float(sums[1])) / B
    assert abs(loss - float(ref)) / abs(float(ref)) < REL
    assert _rel(dT, Tq.grad) < REL and _rel(dV, Vq.grad) < REL
    assert abs(float(rsum) - float(ls.grad)) / abs(float(ls.grad)) < REL
