"""GPU: cluster split-K (K-slices of one output tile on the CTAs of a thread-block cluster, partial tiles folded through
distributed shared memory before the ordinary epilogue; csrc/tic_umma.cuh KC) against an fp64 reference, for every operand
form the small-batch step uses: K-/MN-major operands, split-precision (hi, lo) operands, bias + ReLU, fp32 and bf16(+residual)
outputs, ragged M/N/K.  `tic_gemm_plan` tells which launch shape the dispatcher picks, so the test KNOWS the cluster path ran."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _enable_cluster_splitk(monkeypatch):
    """Cluster split-K is opt-in (its fixed cost only pays for K >= ~2560): the library reads TIC_CLUSTER_K at every call."""
    monkeypatch.setenv("TIC_CLUSTER_K", "96")


def _dev():
    return torch.device("cuda:0")


def _plan(M, N, K, nsplit=0, accumulate=0):
    from tic_b200 import capi
    bn, ks, kc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    capi.call("tic_gemm_plan", M, N, K, nsplit, accumulate, ctypes.byref(bn), ctypes.byref(ks), ctypes.byref(kc))
    return bn.value, ks.value, kc.value


def _padded(t):
    ld = (t.shape[1] + 7) // 8 * 8
    buf = torch.zeros(t.shape[0], ld, dtype=t.dtype, device=t.device)
    buf[:, :t.shape[1]] = t
    return buf, ld


def _split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K,want_kc", [(256, 512, 768, 4), (512, 768, 1536, 2), (200, 136, 520, 2), (128, 64, 1024, 4),
                                           (256, 776, 512, 2)])
def test_cluster_splitk_forms(a_mn, b_mn, M, N, K, want_kc):
    import tic_b200.plan as P
    bn, ks, kc = _plan(M, N, K)
    assert ks == 1 and kc == want_kc, (bn, ks, kc)
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((K, M) if a_mn else (M, K), generator=g).to(torch.bfloat16)
    Bm = torch.randn((K, N) if b_mn else (N, K), generator=g).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    Ap, lda = _padded(A.to(_dev()))
    Bp, ldb = _padded(Bm.to(_dev()))
    D = torch.full((M, N), float("nan"), device=_dev())
    P.gemm(Ap, lda, a_mn, Bp, ldb, b_mn, D, N, 0, M, N, K, alpha=0.5, bias=bias.to(_dev()), relu=True)
    torch.cuda.synchronize()
    Ar, Br = A.double(), Bm.double()
    ref = torch.relu(0.5 * ((Ar.t() if a_mn else Ar) @ (Br if b_mn else Br.t())) + bias.double())
    err = float((D.double().cpu() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err


@pytest.mark.parametrize("M,N,K", [(512, 768, 768), (256, 768, 512), (256, 512, 256)])
def test_cluster_splitk_split_operands_and_bf16_out(M, N, K):
    """(hi, lo) operand pairs add K-segments (the dX / d_t_pool / ITC-gradient GEMMs of the c2 step); bf16 output with residual."""
    import tic_b200.plan as P
    bn, ks, kc = _plan(M, N, K, nsplit=2)
    assert kc >= 2, (bn, ks, kc)
    g = torch.Generator().manual_seed(K)
    A32 = torch.randn(M, K, generator=g)
    B32 = torch.randn(K, N, generator=g)          # MN-major B: row-major [K, N]
    Ah, Al = _split(A32)
    Bh, Bl = _split(B32)
    dev = _dev()
    Ah, Al, Bh, Bl = Ah.to(dev), Al.to(dev), Bh.to(dev), Bl.to(dev)
    D = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    D_lo = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    P.gemm(Ah, K, 0, Bh, N, 1, D, N, 1, M, N, K, A_lo=Al, B_lo=Bl, D_lo=D_lo)
    torch.cuda.synchronize()
    A64 = Ah.double().cpu() + Al.double().cpu()
    B64 = Bh.double().cpu() + Bl.double().cpu()
    # the kernel computes A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (the lo*lo term, ~2^-18 relative, is dropped by design)
    ref = A64 @ B64 - Al.double().cpu() @ Bl.double().cpu()
    got = D.double().cpu() + D_lo.double().cpu()
    err = float((got - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err     # hi+lo bf16 output carries ~16 mantissa bits


def test_cluster_splitk_is_skipped_where_it_does_not_apply():
    assert _plan(4096, 4096, 4096)[2] == 1            # large: fills the machine by itself
    assert _plan(512, 1536, 512, accumulate=1)[2] == 1  # fp32-atomic split-K problems keep their own path
    assert _plan(128, 64, 64)[2] == 1                 # one k-block: nothing to split


@pytest.mark.parametrize("B,Pd", [(256, 512), (128, 768), (200, 256)])
def test_itc_small_batch_uses_cluster_splitk_and_matches_oracle(B, Pd):
    """The narrow-tile similarity kernels (forward softmax epilogue, backward gradient-operand epilogue) behind the same reduction."""
    from oracle import restatement as R
    import math
    import tic_b200.plan as P
    dev = _dev()
    g = torch.Generator().manual_seed(B + Pd)
    T = torch.randn(B, Pd, generator=g).to(torch.bfloat16)
    V = (torch.randn(B, Pd, generator=g) + 0.3 * T.float()).to(torch.bfloat16)
    it = P.ItcPlan(B, B, Pd, dev)
    sums, rsum = torch.zeros(2, device=dev), torch.zeros(1, device=dev)
    dT, dV = torch.empty(B, Pd, device=dev), torch.empty(B, Pd, device=dev)
    ls = 2.6592
    it.run(T.to(dev), V.to(dev), math.exp(ls), 1.0, sums, rsum, dT_f32=dT, dV_f32=dV)
    torch.cuda.synchronize()
    Tr = T.double().requires_grad_(True)
    Vr = V.double().requires_grad_(True)
    l = torch.tensor(ls, dtype=torch.float64, requires_grad=True)
    loss = R.clip_loss(R.itc_logits(Tr, Vr, l))
    loss.backward()
    got = 0.5 * float(sums.sum()) / B
    assert abs(got - float(loss)) / abs(float(loss)) < 1e-3
    for a, b in ((dT, Tr.grad), (dV, Vr.grad)):
        assert float((a.double().cpu() - b).abs().max() / b.abs().max()) < 1e-3
    assert abs(float(rsum) - float(l.grad)) / max(abs(float(l.grad)), 1e-6) < 1e-3
