"""Static-shape execution plans for the late-fusion head + auxiliary-loss step (forward AND backward).

A plan owns every workspace, so a step allocates nothing, performs no host synchronisation and is CUDA-graph
capturable.  All arithmetic happens in libtic_b200.so (hand-written sm_100a kernels, see csrc/); torch is used only
for device memory and streams.

  ItcPlan   image-text contrastive loss on embeddings: row norms -> tcgen05 similarity tiles with the bidirectional
            softmax-CE fused in the epilogue -> tile recompute emitting bf16 gradient operands -> two tcgen05 GEMMs
            -> normalise-backward.  (HF VisionTextDualEncoderModel.forward :268-273 + models/utils.py:225-231.)
  HeadPlan  everything models/mm_late.py does after the encoders return (MM_Model.forward :155-193, the loss mix
            :473-487) for fusion in {concat, attention, gmu, aspect-att}, with ITM sampling/gather (:389-414).
"""
import ctypes
import math
import os
from typing import Dict, Optional

import torch

from . import capi
from .capi import call, ptr

BF16, F32 = torch.bfloat16, torch.float32


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _up8(x):
    return (x + 7) // 8 * 8


def _addr(t):
    return None if t is None else (t if isinstance(t, int) else t.data_ptr())


def gemm(A, lda, a_mn, Bm, ldb, b_mn, D, ldd, d_bf16, M, N, K, alpha=1.0, bias=None, relu=False, A_lo=None, B_lo=None,
         D_lo=None, accumulate=False, row_ss=None):
    """D[M,N] = alpha * op(A) op(B)^T (+bias)(relu) on tcgen05 (tic_gemm_bf16).  A/Bm/D are tensors or raw device
    addresses; A_lo / B_lo make that operand a split-precision bf16 (hi, lo) pair, D_lo receives the bf16 residual.
    row_ss: fp32 [tic_gemm_rowss_parts(N), M] receiving per-tile row sums of squares (fused L2-norm statistics)."""
    if row_ss is not None:
        call("tic_gemm_bf16_rowss", _addr(A), _addr(A_lo), lda, int(a_mn), _addr(Bm), _addr(B_lo), ldb, int(b_mn), _addr(D),
             _addr(D_lo), ldd, int(d_bf16), M, N, K, float(alpha), ptr(bias), int(relu), ptr(row_ss), _stream())
        return
    call("tic_gemm_bf16", _addr(A), _addr(A_lo), lda, int(a_mn), _addr(Bm), _addr(B_lo), ldb, int(b_mn), _addr(D),
         _addr(D_lo), ldd, int(d_bf16), M, N, K, float(alpha), ptr(bias), int(relu), int(accumulate), _stream())


class _Branches:
    """Fork/join helper: `with br(k):` issues the enclosed launches on side stream k (after everything issued so far on
    the current stream); `br.join(k)` makes the current stream wait for it.  Under CUDA-graph capture the branches become
    parallel paths of the graph — the small-batch step is a chain of latency-bound kernels, so width matters."""

    def __init__(self, device, enabled=True, priority=None):
        self.dev, self.enabled, self.side = device, enabled, {}
        self.priority = priority or {}     # branch key -> CUDA stream priority (-1 = high; default 0 = low)
        self.pdl_chains = False

    class _Ctx:
        def __init__(self, outer, k):
            self.o, self.k = outer, k

        def __enter__(self):
            o = self.o
            if not o.enabled:
                return
            if self.k not in o.side:
                o.side[self.k] = torch.cuda.Stream(device=o.dev, priority=o.priority.get(self.k, 0))
            s = o.side[self.k]
            s.wait_stream(torch.cuda.current_stream())
            self.cm = torch.cuda.stream(s)
            self.cm.__enter__()
            # programmatic dependent launch follows the latency-critical chains only (HeadPlan.pdl_chains): a side branch
            # launches its kernels in ordinary stream order, a high-priority branch inherits the chain's mode
            self.prev_pdl = None
            if o.pdl_chains:
                self.prev_pdl = capi.load().tic_set_pdl(1 if o.priority.get(self.k, 0) < 0 else 0)

        def __exit__(self, *a):
            if self.o.enabled:
                if self.prev_pdl is not None:
                    capi.load().tic_set_pdl(self.prev_pdl)
                self.cm.__exit__(*a)

    def __call__(self, k):
        return self._Ctx(self, k)

    def join(self, k):
        if self.enabled and k in self.side:
            torch.cuda.current_stream().wait_stream(self.side[k])


class ItcPlan:
    """ITC forward+backward for a row block of `m` text rows against `n` gathered image columns (single GPU: m == n)."""

    def __init__(self, m: int, n: int, P: int, device, row_offset: int = 0, materialize_logits: bool = False,
                 need_dv: bool = True, precise: Optional[bool] = None, col_sums: bool = True, splitk_grad: Optional[bool] = None,
                 hard: bool = False):
        assert P % 8 == 0, "embedding width must be a multiple of 8 (16-byte TMA rows)"
        self.m, self.n, self.P, self.row_offset = m, n, P, row_offset
        # device scalar exp(logit_scale) (HeadPlan: written by tic_refresh_weights at the head of the step); None = the host
        # floats passed to the methods below are used
        self.scale_dev = None
        # Split-precision gradient operands (bf16 hi+lo): with few negatives the bf16 rounding of the softmax
        # probabilities does not average out; from ~4k columns on it does and the second K-segment is skipped.
        self.precise = (n < 4096) if precise is None else bool(precise)
        self.nrp = capi.load().tic_itc_row_parts(n)
        self.ncp = capi.load().tic_itc_col_parts(m)
        dev = device
        self.rinv_t = torch.empty(m, dtype=F32, device=dev)
        self.rinv_v = torch.empty(n, dtype=F32, device=dev)
        self.row_part = torch.empty(self.nrp, m, dtype=F32, device=dev)
        # col_sums=False (symmetric peer-memory mode): column statistics are the row statistics of the swapped block
        self.col_part = torch.empty(self.ncp, n, dtype=F32, device=dev) if col_sums else None
        self.col_sum = torch.empty(n, dtype=F32, device=dev) if col_sums else None   # NCCL path (all-reduced)
        self.diag = torch.empty(m, dtype=F32, device=dev)
        self.lse_row = torch.empty(m, dtype=F32, device=dev)
        self.lse_col = torch.empty(n, dtype=F32, device=dev)
        self.ld_ga, self.ld_gbt = _up8(n), _up8(m)
        self.GA = torch.empty(m, self.ld_ga, dtype=BF16, device=dev)
        # precise (small batch): separate transposed operand GBT with residuals.  Otherwise "GA-shared": the image-side
        # gradient GEMM reads GA MN-major against a normalised bf16 copy of the text embeddings (no GBT at all).
        self.shared_ga = need_dv and not self.precise
        # "GB": Gp * rinv_t, row-major with GA's layout [m, ld_ga]; the image-side GEMM reads it MN-major (no transposed operand)
        self.GBT = torch.empty(m, self.ld_ga, dtype=BF16, device=dev) if (need_dv and self.precise) else None
        self.GA_lo = torch.empty(m, self.ld_ga, dtype=BF16, device=dev) if self.precise else None
        self.GBT_lo = torch.empty(m, self.ld_ga, dtype=BF16, device=dev) if (need_dv and self.precise) else None
        self.That = torch.empty(m, P, dtype=BF16, device=dev) if self.shared_ga else None
        self.acc_t = torch.empty(m, P, dtype=F32, device=dev)
        self.acc_v = torch.empty(n, P, dtype=F32, device=dev) if need_dv else None
        self.logits = torch.empty(m, n, dtype=F32, device=dev) if materialize_logits else None
        self.need_dv = need_dv
        # hard-negative sampling in tile-stream form (a-5'): per (row, column part) integer weight sums + the located part
        self.nqp = capi.load().tic_itc_q_parts(n)
        self.qpart = torch.empty(self.nqp, m, dtype=torch.int64, device=dev) if hard else None
        self.loc_part = torch.empty(m, dtype=torch.int32, device=dev) if hard else None
        self.loc_res = torch.empty(m, dtype=torch.int64, device=dev) if hard else None
        # dT = GA[m,n] V[n,P] with m << n (multi-GPU row block at small per-rank batch): the un-split GEMM has m/128 * P/64
        # CTAs walking the whole K = n; fp32-atomic split-K fills the machine instead (nondeterministic summation order)
        self.splitk_grad = (m <= 1024 and n >= 4 * m) if splitk_grad is None else bool(splitk_grad)

    # -- pieces (the distributed path interleaves collectives between them) --
    def norm_t(self, T, ldt, T_lo=None):
        call("tic_row_rnorm_bf16", ptr(T), ptr(T_lo), ldt, self.m, self.P, ptr(self.rinv_t), ptr(self.That), self.P, _stream())

    def norm_v(self, V, ldv, V_lo=None):
        call("tic_row_rnorm_bf16", ptr(V), ptr(V_lo), ldv, self.n, self.P, ptr(self.rinv_v), None, 0, _stream())

    def norms(self, T, ldt, V, ldv, t_only=False, T_lo=None, V_lo=None):
        self.norm_t(T, ldt, T_lo=T_lo)
        if not t_only:
            self.norm_v(V, ldv, V_lo=V_lo)

    def fwd_tiles(self, T, ldt, V, ldv, scale, T_lo=None, V_lo=None, ss_t=None, ss_v=None, seg=None):
        """ss_t / ss_v: row sum-of-squares partials from the projection GEMMs (gemm(row_ss=...)); the tiles then compute and
        write rinv_t / rinv_v themselves and norm_t()/norm_v() are not needed.
        seg = (ready, epoch, seg_cols, my_seg): V is being pulled from the peers by a concurrent kernel, segment p of
        seg_cols columns is usable once ready[p] >= epoch[0]; the tiles are visited local-segment-first."""
        sr, se, sc, ms = (ptr(seg[0]), ptr(seg[1]), int(seg[2]), int(seg[3])) if seg is not None else (None, None, 0, 0)
        call("tic_itc_fwd", ptr(T), ptr(T_lo), ldt, ptr(V), ptr(V_lo), ldv, ptr(self.rinv_t), ptr(self.rinv_v), self.m, self.n, self.P,
             self.row_offset, float(scale), float(scale), ptr(self.row_part), ptr(self.col_part), ptr(self.diag),
             ptr(self.logits), self.n if self.logits is not None else 0, ptr(ss_t), 0 if ss_t is None else ss_t.shape[0],
             ptr(ss_v), 0 if ss_v is None else ss_v.shape[0], sr, se, sc, ms, ptr(self.scale_dev), ptr(self.qpart), _stream())

    @property
    def can_fuse_small(self):
        """forward + gradient-operand tiles in ONE cluster launch (at most 8 tiles of 128 x 64; csrc/itc.cu)"""
        return (self.col_part is not None and self.m == self.n and self.row_offset == 0 and
                bool(capi.load().tic_itc_fused_small_ok(self.m, self.n)))

    def fwd_bwd_fused(self, T, ldt, V, ldv, scale, gscale, T_lo=None, V_lo=None, ss_t=None, ss_v=None):
        call("tic_itc_fwd_bwd_small", ptr(T), ptr(T_lo), ldt, ptr(V), ptr(V_lo), ldv, ptr(self.rinv_t), ptr(self.rinv_v), self.m,
             self.n, self.P, self.row_offset, float(scale), ptr(self.row_part), ptr(self.col_part), ptr(self.diag),
             ptr(self.logits), self.n if self.logits is not None else 0, ptr(ss_t), 0 if ss_t is None else ss_t.shape[0],
             ptr(ss_v), 0 if ss_v is None else ss_v.shape[0], ptr(self.scale_dev), ptr(self.qpart), float(gscale), ptr(self.GA),
             self.ld_ga, ptr(self.GBT), self.ld_gbt, ptr(self.GA_lo), ptr(self.GBT_lo), _stream())

    def hard_locate(self, u_coin, u_pick, labels, src_idx):
        """labels / default sources for every row + (part, residual) of each mismatch row from the weight sums of fwd_tiles"""
        call("tic_itm_hard_locate", ptr(u_coin), ptr(u_pick), self.m, self.n, self.row_offset, ptr(self.qpart), self.nqp,
             ptr(labels), ptr(src_idx), ptr(self.loc_part), ptr(self.loc_res), _stream())

    def hard_pick(self, T, ldt, V, ldv, scale, src_idx, T_lo=None, V_lo=None):
        """tile recompute: src_idx of the located rows (same operands and tile shapes as fwd_tiles -> identical logits)"""
        call("tic_itc_pick", ptr(T), ptr(T_lo), ldt, ptr(V), ptr(V_lo), ldv, ptr(self.rinv_t), ptr(self.rinv_v), self.m, self.n,
             self.P, self.row_offset, float(scale), float(scale), ptr(self.scale_dev), ptr(self.loc_part), ptr(self.loc_res),
             ptr(src_idx), _stream())

    def lse_loss(self, scale, loss_sums, col_parts=None, n_col_parts=None):
        cp = self.col_part if col_parts is None else col_parts
        ncp = self.ncp if n_col_parts is None else n_col_parts
        call("tic_itc_lse_loss", ptr(self.row_part), self.nrp, ptr(cp), ncp, ptr(self.diag), self.m, self.n,
             self.row_offset, float(scale), ptr(self.lse_row), ptr(self.lse_col), ptr(loss_sums), ptr(self.scale_dev), _stream())

    def bwd_operands(self, T, ldt, V, ldv, scale, gscale, T_lo=None, V_lo=None, inline_lse=False, seg=None):
        """inline_lse: derive lse_row/lse_col inside the kernel from the forward partials (few partials = small batch), so
        that lse_loss() is only needed for the loss value and can run beside the backward instead of before it."""
        rp = (ptr(self.row_part), self.nrp) if inline_lse else (None, 0)
        cp = (ptr(self.col_part), self.ncp) if inline_lse else (None, 0)
        sr, se, sc, ms = (ptr(seg[0]), ptr(seg[1]), int(seg[2]), int(seg[3])) if seg is not None else (None, None, 0, 0)
        call("tic_itc_bwd_g", ptr(T), ptr(T_lo), ldt, ptr(V), ptr(V_lo), ldv, ptr(self.rinv_t), ptr(self.rinv_v), ptr(self.lse_row),
             ptr(self.lse_col), self.m, self.n, self.P, float(scale), float(gscale), ptr(self.GA), self.ld_ga,
             ptr(self.GBT), self.ld_gbt, ptr(self.GA_lo), ptr(self.GBT_lo), rp[0], rp[1], cp[0], cp[1], float(scale),
             ptr(self.scale_dev), sr, se, sc, ms, _stream())

    @property
    def can_inline_lse(self):
        # every tile re-reads its nrp + ncp partials: free while the step is latency-bound, but from ~32 partials on the load
        # chain outlasts the tile's MMAs (c5 point 8192 x 256: tic_itc_bwd_g 220 us inline vs ~75 us after the 7 us lse kernel;
        # profiles/r02_c5_timelines.txt).  TIC_INLINE_LSE_MAX: A/B switch.
        cap = int(os.environ.get("TIC_INLINE_LSE_MAX", "16"))      # column partials (= row tiles of 128): B <= 2048
        return self.col_part is not None and self.nrp <= 2 * cap and self.ncp <= cap

    def grad_gemm_t(self, V, ldv, V_lo=None, prezeroed=False):
        # dT_acc[m,P] = GA[m,n] * V[n,P]   (A K-major, B = V read MN-major: no transposed copy of V)
        if self.splitk_grad:   # few output tiles, long K (a small row block against many gathered columns): split K over CTAs
            if not prezeroed:  # (the multi-GPU sequencing zeroes acc_t on a side branch beside the gradient-operand tiles)
                self.acc_t.zero_()
            gemm(self.GA, self.ld_ga, 0, V, ldv, 1, self.acc_t, self.P, 0, self.m, self.P, self.n, A_lo=self.GA_lo, B_lo=V_lo,
                 accumulate=True)
            return
        gemm(self.GA, self.ld_ga, 0, V, ldv, 1, self.acc_t, self.P, 0, self.m, self.P, self.n, A_lo=self.GA_lo, B_lo=V_lo)

    def grad_gemm_v(self, T, ldt, T_lo=None):
        if self.shared_ga:
            # dV_acc'[n,P] = GA^T[n,m] * That[m,P]: GA is read MN-major (no transposed operand in HBM); rows carry rinv_v[j]
            gemm(self.GA, self.ld_ga, 1, self.That, self.P, 1, self.acc_v, self.P, 0, self.n, self.P, self.m)
        else:
            # dV_acc[n,P] = GB^T[n,m] * T[m,P]   (GB row-major [m,n]: read MN-major)
            gemm(self.GBT, self.ld_ga, 1, T, ldt, 1, self.acc_v, self.P, 0, self.n, self.P, self.m, A_lo=self.GBT_lo,
                 B_lo=T_lo)

    def grad_gemms(self, T, ldt, V, ldv, T_lo=None, V_lo=None):
        self.grad_gemm_t(V, ldv, V_lo=V_lo)
        if self.need_dv:
            self.grad_gemm_v(T, ldt, T_lo=T_lo)

    def finalize_t(self, T, ldt, V_diag, ldv, rinv_v_diag, scale, diag_coef, dT_f32, dT_bf16, r_sum, dT_lo=None,
                   T_lo=None, V_diag_lo=None):
        call("tic_itc_grad_finalize", ptr(self.acc_t), self.P, ptr(T), ptr(T_lo), ldt, ptr(self.rinv_t), ptr(V_diag),
             ptr(V_diag_lo), ldv,
             ptr(rinv_v_diag), self.m, self.P, float(scale), float(diag_coef), ptr(dT_f32), self.P, ptr(dT_bf16), ptr(dT_lo),
             self.P, ptr(r_sum), 0, ptr(self.scale_dev), _stream())

    def finalize_v(self, acc_v, V, ldv, rinv_v, T_diag, ldt, rinv_t_diag, rows, scale, diag_coef, dV_f32, dV_bf16,
                   dV_lo=None, V_lo=None, T_diag_lo=None):
        call("tic_itc_grad_finalize", ptr(acc_v), self.P, ptr(V), ptr(V_lo), ldv, ptr(rinv_v), ptr(T_diag), ptr(T_diag_lo),
             ldt, ptr(rinv_t_diag),
             rows, self.P, float(scale), float(diag_coef), ptr(dV_f32), self.P, ptr(dV_bf16), ptr(dV_lo), self.P, None,
             1 if self.shared_ga else 0, ptr(self.scale_dev), _stream())

    # -- single-GPU convenience: full forward + backward --
    def run(self, T, V, scale, g, loss_sums, r_sum, dT_f32=None, dT_bf16=None, dV_f32=None, dV_bf16=None):
        """T [m,P], V [n,P] bf16 (m == n, row_offset == 0).  g = dLoss/d(clip_loss).  loss_sums[2], r_sum[1] must be
        zero on entry: loss_sums -> (sum_i lse_row-diag, sum_i lse_col-diag), r_sum -> dLoss/dlogit_scale."""
        assert self.m == self.n and self.row_offset == 0
        ldt, ldv = T.stride(0), V.stride(0)
        self.norms(T, ldt, V, ldv)
        self.fwd_tiles(T, ldt, V, ldv, scale)
        self.lse_loss(scale, loss_sums)
        self.bwd_operands(T, ldt, V, ldv, scale, g / (2.0 * self.n))
        self.grad_gemms(T, ldt, V, ldv)
        self.finalize_t(T, ldt, V, ldv, self.rinv_v, scale, g / self.n, dT_f32, dT_bf16, r_sum)
        if self.need_dv:
            self.finalize_v(self.acc_v, V, ldv, self.rinv_v, T, ldt, self.rinv_t, self.n, scale, g / self.n, dV_f32, dV_bf16)


class HeadPlan:
    """One training step of the late-fusion head on given encoder outputs: forward, losses, backward."""

    FUSIONS = ("concat", "attention", "gmu", "aspect-att", None)

    def __init__(self, B: int, *, E: int = 768, P: Optional[int] = 512, C: int = 4, fusion: Optional[str] = "concat",
                 use_itc: bool = True, use_itm: bool = True, beta_itc: float = 0.1, beta_itm: float = 0.1, Lv: int = 197,
                 itm_mode: str = "uniform", materialize_logits: bool = False, device="cuda",
                 split_precision: Optional[bool] = None):
        if fusion not in self.FUSIONS:
            raise KeyError("fusion_name %r is not implemented for ViT-family encoders (mm_late.py:92-144 implements "
                           "concat, attention, aspect-att, gmu; xatt/concat_cnn resolve to undefined names :44-45)" % fusion)
        if fusion == "aspect-att" and use_itm:
            # mm_late.py:181 calls mm_fusion without pools on the ITM branch -> torch.stack((None, None)) raises TypeError
            raise TypeError("aspect-att cannot be combined with the ITM loss (the reference raises TypeError at mm_late.py:181)")
        capi.load()
        self.B, self.E, self.P, self.C, self.fusion = B, E, P, C, fusion
        self.use_itc, self.use_itm = use_itc, use_itm and fusion is not None
        self.beta_itc, self.beta_itm = (beta_itc if use_itc else 0.0), (beta_itm if self.use_itm else 0.0)
        self.Lv, self.itm_mode = Lv, {"uniform": 0, "hard": 1}[itm_mode]
        self.dev = torch.device(device)
        self.R = 2 * B if self.use_itm else B
        # Split-precision (bf16 hi+lo) intermediates in the fusion chain keep every gradient within 1e-3 (max-norm) of the
        # fp64 oracle.  split_precision=False is an opt-in fast mode: plain bf16 intermediates, 2-3x less tensor work in
        # the fusion GEMMs (c4: 1.61 -> 1.34 ms), gradients within 5e-3 (measured 1.7e-3 .. 3.2e-3: the rounding error of a
        # zero-mean sum does not average out relative to the sum) — NOT the parity-validated default.
        self.split = True if split_precision is None else bool(split_precision)
        self.Pe = P if P is not None else E  # width of the contrastive embeddings
        if fusion is None:
            self.w_cls = 0.0
            self.g_itc = 1.0
        else:
            self.w_cls = 1.0 - (self.beta_itc + self.beta_itm)
            self.g_itc = self.beta_itc
        if self.itm_mode == 1 and not (use_itc and self.use_itm):
            raise ValueError("itm_mode='hard' samples negatives from the ITC similarities: it needs use_itc and use_itm")
        self.itc = ItcPlan(B, B, self.Pe, self.dev, materialize_logits=materialize_logits,
                           hard=self.itm_mode == 1) if use_itc else None
        self._alloc()
        self.w: Dict[str, torch.Tensor] = {}
        self.scale = math.exp(2.6592)            # host copy of exp(logit_scale): informational once weights are bound
        # exp(logit_scale) as a DEVICE scalar: every ITC kernel of this plan reads it through a pointer, so a captured step
        # follows a trainable logit_scale (mm_late.py:59-69); scale_status != 0 if it ever left (0, 40]
        self.scale_t = torch.full((1,), self.scale, dtype=F32, device=self.dev)
        self.scale_status = torch.zeros(1, dtype=torch.int32, device=self.dev)
        if self.itc is not None:
            self.itc.scale_dev = self.scale_t
        import os as _os2
        self.fuse_itc_small = _os2.environ.get("TIC_ITC_FUSED_SMALL", "1") != "0"     # A/B switch (same kernels' epilogues)
        # code warm-up launches at the head of the step (see _warm_calls): only where the step is latency-bound
        self.code_warm = (_os2.environ.get("TIC_CODE_WARM", "1") != "0") and B <= 512
        self._warm = None
        E_ = self.E
        self._warm_H = torch.zeros(2, E_, dtype=F32, device=self.dev)
        self._warm_P = torch.zeros(2, E_, dtype=F32, device=self.dev)
        self._warm_dH = torch.zeros(2, E_, dtype=BF16, device=self.dev)
        self._warm_dH_lo = torch.zeros(2, E_, dtype=BF16, device=self.dev)
        self._warm_dP = [torch.zeros(2, E_, dtype=BF16, device=self.dev) for _ in range(4)]
        self.live_weights = False                # bind_params(live=True): the bf16 working copies are refreshed inside step()
        self._refresh_groups = {}
        self.generation = 0                      # autograd mode: bumped by forward(); backward() checks it (see mm_late._HeadFn)
        self.parallel_streams = True
        import os as _os
        # stream priorities only pay while the step is latency-bound (small batch); with persistent 148-CTA kernels they
        # starve the concurrent side work instead (measured: c3 on 2 GPUs 0.331 -> 0.361 ms).  TIC_HI_PRIORITY=0/1 overrides.
        _hp = _os.environ.get("TIC_HI_PRIORITY")
        self.hi_priority_chains = (B <= 1024) if _hp is None else (_hp != "0")
        self._side = None
        self._hi0 = None
        self._ev_heads = None
        # programmatic dependent launch on the kernels of the latency-critical chains only (TIC_PDL_CHAINS=1; needs the
        # high-priority-chain mode): see _Branches and tic_set_pdl
        # Round 2: ON by default where the chains run on high-priority streams (B <= 1024).  With two CTAs per SM for the small
        # tcgen05 launches and the heads kernel confined to part of the machine, the chains no longer starve each other and PDL
        # pays on the full step too (c2: 73.2 -> 66.1 us; profiles/r02_c2_ab.txt).  TIC_PDL_CHAINS=0 turns it off.
        _pc = _os.environ.get("TIC_PDL_CHAINS")
        # (single-GPU plan only: the multi-GPU plans run spinning exchange kernels beside their tile kernels, where early-
        # launched waiting CTAs are a liability)
        self.pdl_chains = (self.hi_priority_chains and type(self).__name__ == "HeadPlan") if _pc is None else (_pc == "1")
        # TIC_STEP_TAIL bit 0: small memset on a side branch, bit 1: loss mix issued before the backward (A/B switch).
        # Measured on the c2 graph (scripts/timeline.py --plain-only, 400 replays): 0 -> 92.2 us, 1 -> 96.4 us (a memset node
        # joined into both chains costs more than the ~2 us it takes at the head), 2 -> 90.3 us, 3 -> 92.7 us.  Default 2.
        _tail = int(_os.environ.get("TIC_STEP_TAIL", "2"))
        self._defer_small_zero, self._early_mix = bool(_tail & 1), bool(_tail & 2)
        self.br = _Branches(self.dev, enabled=True, priority={"v": -1, "cb": -1} if self.hi_priority_chains else None)

    # ------------------------------------------------------------------ buffers
    def _alloc(self):
        B, E, C, R, dev = self.B, self.E, self.C, self.R, self.dev
        e = lambda *s, dt=F32: torch.empty(*s, dtype=dt, device=dev)  # noqa: E731
        # one contiguous block of accumulators that must be zero at the start of every step (one memset)
        sizes = {"losses": 2, "itc_sums": 2, "r_sum": 1, "_pad": 3, "dW_cls": C * E, "db_cls": C, "dW_tim": 2 * E,
                 "db_tim": 2, "db_f": E, "db_Q": E, "db_V": E, "db_gt": 2 * E, "db_gv": 2 * E, "dw_a": E, "db_a": 1}
        n_small_keys = len(sizes)
        # Weight gradients: a GEMM the library would split over K (long K, few output tiles: large batches) accumulates with
        # fp32 atomics and must start at zero; one it runs un-split (the small-batch step: K = B is short) writes plain
        # stores — no atomics, nothing to zero: at c2 the 8.5 MB memset of round 1 is gone.  (M, N, K, split operands):
        import os as _os
        sp = 1 if self.split else 0
        self.pairwise = self.fusion == "concat" and _os.environ.get("TIC_CONCAT_PAIRWISE", "1") != "0"   # A/B switch
        big = {}
        if self.P is not None:
            big.update(dW_t=(self.P * E, (self.P, E, B, sp)), dW_v=(self.P * E, (self.P, E, B, sp)))
        if self.fusion in ("concat", "attention", "gmu"):
            if self.pairwise:   # two [E, E] halves, K = B; d_xt_cls is written by its GEMM directly
                big.update(dW_f=(E * 2 * E, (E, E, B, sp)), d_xt_cls=(B * E, None))
            else:               # d_xt_cls: accumulated by tic_unpack_cls_grad (atomics)
                big.update(dW_f=(E * 2 * E, (E, 2 * E, R, sp + (0 if self.fusion == "concat" else sp))), d_xt_cls=(B * E, "atomic"))
        if self.fusion == "attention":
            big.update(dW_Q=(E * E, (E, E, R, sp)), dW_V=(E * E, (E, E, R, 2 * sp)), dWK_aug=(E * (E + 8), (E, E + 8, R, 2 * sp)))
        if self.fusion == "gmu":
            big.update(dW_gt=(2 * E * E, (2 * E, E, R, sp)), dW_gv=(2 * E * E, (2 * E, E, R, sp)))
        self._atomic = {}
        tn, ks, kc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        for k, (n, shape) in big.items():
            if shape is None:
                self._atomic[k] = False
            elif shape == "atomic":
                self._atomic[k] = True
            else:
                call("tic_gemm_plan", shape[0], shape[1], shape[2], shape[3], 1, ctypes.byref(tn), ctypes.byref(ks), ctypes.byref(kc))
                self._atomic[k] = ks.value > 1
        plain = {k: n for k, (n, _) in big.items() if not self._atomic[k]}
        sizes.update({k: n for k, (n, _) in big.items() if self._atomic[k]})
        pad = (-sum(sizes.values())) % 4
        sizes["_pad2"] = pad
        self.zb = torch.zeros(sum(sizes.values()), dtype=F32, device=dev)
        self.pb = torch.zeros(max(sum(plain.values()), 1), dtype=F32, device=dev)     # overwritten by plain stores every step
        self.z, off = {}, 0
        for k, n in sizes.items():
            self.z[k] = self.zb[off:off + n]
            off += n
        off = 0
        for k, n in plain.items():
            self.z[k] = self.pb[off:off + n]
            off += n
        # the first keys (loss sums, small-head gradients) are needed right away; the large weight-gradient accumulators only
        # by the parameter-gradient branches and the final unpack: their memset runs on a side branch (_zero_accumulators)
        n_small = sum(n for k, n in list(sizes.items())[:n_small_keys])
        self.zb_small, self.zb_big = self.zb[:n_small], self.zb[n_small:]
        self.out: Dict[str, torch.Tensor] = {"loss": e(4)}
        o = self.out
        if self.P is not None:
            self.Y = e(2 * B, self.P, dt=BF16)       # [Yt ; Yv] projected embeddings (split hi/lo: produced on the device)
            self.Y_lo = e(2 * B, self.P, dt=BF16)
            self.dY = e(2 * B, self.P, dt=BF16)      # gradients w.r.t. Y (GEMM operands, split hi/lo)
            self.dY_lo = e(2 * B, self.P, dt=BF16)
            o["dW_t"], o["dW_v"] = self.z["dW_t"].view(self.P, E), self.z["dW_v"].view(self.P, E)
            o["d_t_pool"] = e(B, E)
        else:
            o["d_t_emb"], o["d_v_emb"] = e(B, self.Pe), e(B, self.Pe)
        o["d_logit_scale"] = self.z["r_sum"]
        if self.fusion is None:
            return
        o["out_cls"], o["lbl_tim"], o["src_idx"] = e(B, C), torch.empty(B, dtype=torch.int64, device=dev), \
            torch.empty(B, dtype=torch.int32, device=dev)
        if self.use_itm:
            o["out_tim"] = e(B, 2)
        self.H = e(R, E)                              # post-ReLU fusion output (mm_features = H[:B])
        o["mm_features"] = self.H[:B]
        self.dHb, self.dHb_lo = e(R, E, dt=BF16), e(R, E, dt=BF16)
        self.heads_ws = e(R * 8)
        self.dz_ext = e(R, 8)                          # autograd mode: upstream dL/dlogits, padded to 8 columns
        self.y_dummy = torch.zeros(B, C, dtype=F32, device=dev)
        o["dW_cls"], o["db_cls"] = self.z["dW_cls"].view(C, E), self.z["db_cls"]
        o["dW_tim"], o["db_tim"] = self.z["dW_tim"].view(2, E), self.z["db_tim"]
        self.dHf = None
        if self.fusion == "aspect-att":
            self.alpha = e(B, 2)
            self.dHf = e(B, E)
            o["d_t_pool_fusion"] = e(B, E)
            o["dw_a"], o["db_a"] = self.z["dw_a"], self.z["db_a"]
            return
        if self.pairwise:
            # concat in pairwise form (csrc/heads.cu): the two halves of linear_fusion projected once per sample
            self.Pt, self.Pv = e(B, E), e(B, E)
            self.dPt, self.dPt_lo, self.dPv, self.dPv_lo = (e(B, E, dt=BF16) for _ in range(4))
        else:
            self.Xcat = e(R, 2 * E, dt=BF16)
            self.dXt = e(R, E)                            # gradient of the text half of Xcat (fp32)
        o["d_xt_cls"] = self.z["d_xt_cls"].view(B, E)
        o["dW_f"], o["db_f"] = self.z["dW_f"].view(E, 2 * E), self.z["db_f"]
        if self.fusion == "attention":
            Ea = E + 8
            self.Xcat_lo = torch.zeros(R, 2 * E, dtype=BF16, device=dev)   # text half stays 0 (inputs are exact)
            self.q0, self.q0_lo, self.kq = e(R, E, dt=BF16), e(R, E, dt=BF16), e(R, Ea)
            self.xbar_b, self.xbar_lo, self.xbar_f, self.attn = e(R, E, dt=BF16), e(R, E, dt=BF16), e(R, E), e(R, self.Lv)
            self.dctx, self.dctx_lo, self.dxbar = e(R, E, dt=BF16), e(R, E, dt=BF16), e(R, E)
            self.dkq, self.dkq_lo = e(R, Ea, dt=BF16), e(R, Ea, dt=BF16)
            self.dq0, self.dq0_lo, self.dXt2 = e(R, E, dt=BF16), e(R, E, dt=BF16), e(R, E)
            o["dW_Q"], o["db_Q"] = self.z["dW_Q"].view(E, E), self.z["db_Q"]
            o["dW_V"], o["db_V"] = self.z["dW_V"].view(E, E), self.z["db_V"]
            self.dWK_aug = self.z["dWK_aug"].view(E, Ea)
            o["dW_K"], o["db_K"] = self.dWK_aug[:, :E], self.dWK_aug[:, E]
        if self.fusion == "gmu":
            self.tp, self.vp = e(R, 2 * E), e(R, 2 * E)
            self.G, self.G_lo, self.dG = e(R, 2 * E, dt=BF16), e(R, 2 * E, dt=BF16), e(R, 2 * E)
            self.dtp, self.dvp, self.dXg, self.dXt2 = e(R, 2 * E, dt=BF16), e(R, 2 * E, dt=BF16), e(R, 2 * E), e(R, E)
            self.dtp_lo, self.dvp_lo = e(R, 2 * E, dt=BF16), e(R, 2 * E, dt=BF16)
            o["dW_gt"], o["db_gt"] = self.z["dW_gt"].view(2 * E, E), self.z["db_gt"]
            o["dW_gv"], o["db_gv"] = self.z["dW_gv"].view(2 * E, E), self.z["db_gv"]

    # ------------------------------------------------------------------ weights
    def bind_params(self, p: Dict[str, torch.Tensor], live: bool = True):
        """p: fp32 tensors under the reference's state-dict names (mm_late.py:59-89).  The plan keeps POINTERS to them: the
        small fp32 layers are read in place, the large matrices get bf16 working copies that ONE multi-tensor kernel per chain
        (tic_refresh_weights) rewrites from the masters — together with exp(logit_scale) — at the head of every step() when
        live=True.  An optimiser that updates the parameters in place between replays of a captured step is therefore seen
        by the next replay: the captured step is a training step.  Tensors that are not fp32 / contiguous / on the plan's
        device are copied once (and are then NOT live: rebind after changing them).  No host synchronisation."""
        E, dev = self.E, self.dev
        w = self.w
        self._masters = {}

        def master(src):
            return src.detach().to(device=dev, dtype=F32).contiguous()

        groups = {"itc": [], "fusion": []}

        def cast(group, name, src, cols_pad=None, dst_col=0):
            m = master(src)
            if m.dim() == 1:
                m = m.reshape(-1, 1)
            rows, cols = m.shape
            if dst_col == 0:
                ld = cols if cols_pad is None else cols_pad
                if name not in w or w[name].shape != (rows, ld):
                    w[name] = torch.zeros(rows, ld, dtype=BF16, device=dev)
            dst = w[name]
            self._masters[name + "@%d" % dst_col] = m
            groups[group].append((m.data_ptr(), dst.data_ptr() + 2 * dst_col, m.stride(0), dst.stride(0), rows, cols))

        def keep(name, src):
            w[name] = master(src)

        self._logit_scale = master(p["dual_encoder.logit_scale"]).reshape(1)
        if self.P is not None:
            cast("itc", "W_t", p["dual_encoder.text_projection.weight"])
            cast("itc", "W_v", p["dual_encoder.visual_projection.weight"])
        if self.fusion is not None:
            keep("W_cls", p["linear_cls.weight"]); keep("b_cls", p["linear_cls.bias"])
            keep("W_tim", p["linear_tim.weight"]); keep("b_tim", p["linear_tim.bias"])
            if self.fusion == "aspect-att":
                keep("w_a", p["aspectattention.weight"].reshape(-1)); keep("b_a", p["aspectattention.bias"].reshape(-1))
            else:
                cast("fusion", "W_f", p["linear_fusion.weight"]); keep("b_f", p["linear_fusion.bias"])
            if self.fusion == "attention":
                cast("fusion", "W_Q", p["fc_Q.weight"]); keep("b_Q", p["fc_Q.bias"])
                cast("fusion", "W_V", p["fc_V.weight"]); keep("b_V", p["fc_V.bias"])
                cast("fusion", "W_Kaug", p["fc_K.weight"], cols_pad=E + 8)  # [W_K | b_K | 0..]: column E carries <q0, b_K>
                cast("fusion", "W_Kaug", p["fc_K.bias"], dst_col=E)
            if self.fusion == "gmu":
                cast("fusion", "W_gt", p["linear_gmu_t.weight"]); keep("b_gt", p["linear_gmu_t.bias"])
                cast("fusion", "W_gv", p["linear_gmu_v.weight"]); keep("b_gv", p["linear_gmu_v.bias"])
        self._refresh_groups = {}
        if len(groups["itc"]) + len(groups["fusion"]) <= 8:   # ONE launch for both chains (the root node of the captured step)
            groups["all"] = groups["itc"] + groups["fusion"]
        for g, items in groups.items():
            n = len(items)
            if n == 0 and g != "itc":
                continue
            self._refresh_groups[g] = (n, (ctypes.c_void_p * max(n, 1))(*[i[0] for i in items]),
                                       (ctypes.c_void_p * max(n, 1))(*[i[1] for i in items]),
                                       (ctypes.c_int64 * max(n, 1))(*[i[2] for i in items]),
                                       (ctypes.c_int64 * max(n, 1))(*[i[3] for i in items]),
                                       (ctypes.c_int * max(n, 1))(*[i[4] for i in items]),
                                       (ctypes.c_int * max(n, 1))(*[i[5] for i in items]))
        self.live_weights = bool(live)
        self._refresh("itc", force=True)
        self._refresh("fusion", force=True)

    def _refresh(self, group, force=False):
        """fp32 masters -> bf16 working copies of one chain's matrices (+ exp(logit_scale) with the ITC group): one launch"""
        if not (force or self.live_weights) or group not in self._refresh_groups:
            return
        if not force and group != "all" and getattr(self, "_refreshed_all", False):
            return                   # the step's root launch refreshed both chains' matrices already
        n, src, dst, lds, ldd, rows, cols = self._refresh_groups[group]
        with_scale = group in ("itc", "all")
        # live mode inside a step: this launch also zeroes the chain's share of the small accumulator block, so no memset
        # node sits in front of the first kernel of either chain (see _zero_accumulators)
        z0 = z1 = None
        if not force and self._refresh_zeroes:
            zs, n_s = self.zb_small, self.zb_small.numel()
            if group == "all":
                z0 = zs
            elif group == "itc":     # itc_sums, r_sum (+ pad): floats [2, 8)
                z0 = zs[2:8] if "fusion" in self._refresh_groups else zs
            else:                    # losses [0, 2) and the small heads' gradient accumulators [8, n_small)
                z0, z1 = (zs[0:2], zs[8:n_s]) if ("itc" in self._refresh_groups and self.use_itc) else (zs, None)
        call("tic_refresh_weights", n, src, dst, lds, ldd, rows, cols, ptr(self._logit_scale) if with_scale else None,
             ptr(self.scale_t) if with_scale else None, ptr(self.scale_status) if with_scale else None,
             ptr(z0), 0 if z0 is None else z0.numel(), ptr(z1), 0 if z1 is None else z1.numel(), _stream())

    def set_weights(self, p: Dict[str, torch.Tensor]):
        """Snapshot form of bind_params: casts once, now; step() does not refresh.  Also reads exp(logit_scale) back into
        `self.scale` (one host synchronisation, outside any step) for callers that want the value on the host."""
        self.bind_params(p, live=False)
        self.scale = float(self.scale_t.item())

    # ------------------------------------------------------------------ phases
    # fused mode   : step()                      = itc_fwd, fusion_fwd, heads (losses fused), loss mix, fusion_bwd, itc_bwd
    # autograd mode: forward() + backward(grads) = the same kernels, the losses live in the caller (reference train loop)
    def _itc_operands(self, inp):
        B, E, w = self.B, self.E, self.w
        if self.P is not None:
            tp_, vp_ = inp["t_pool"], inp["v_pool"]
            # Embeddings projected on the device are fp32 values: the similarity tiles always consume them as bf16 (hi, lo) pairs
            # (a single-bf16 embedding moves every logit by ~1.5e-3 and the ITC gradients by 1.7e-3 at B = 4096 — over the 1e-3
            # bar; measured).  Only the gradient GEMMs drop the residual K-segments at large batch (_itc_gemm_lo).
            return self.Y[:B], self.Y[B:], self.Y_lo[:B], self.Y_lo[B:]
        return inp["t_pool"], inp["v_pool"], None, None

    def _itc_gemm_lo(self, lo):
        """residual of an embedding as a GEMM operand: kept while the step is latency-bound (< 4096 negatives), dropped where
        tensor time matters (the rounding of a GEMM operand averages out over >= 4096 terms; measured < 1e-3)"""
        return lo if self.itc.precise else None

    def _itc_fwd(self, inp, with_loss=True):
        B, E, w, it = self.B, self.E, self.w, self.itc
        Yt, Yv, Ytl, Yvl = self._itc_operands(inp)
        ldt, ldv = Yt.stride(0), Yv.stride(0)
        br = self.br
        br.enabled = self.parallel_streams
        self._refresh("itc")                              # live weights: W_t / W_v working copies + exp(logit_scale)
        fused_norm = self.P is not None and B <= 1024     # small batch: norm statistics ride on the projection GEMMs
        if fused_norm and getattr(self, "ss_t", None) is None:
            nss = capi.load().tic_gemm_rowss_parts(self.P)
            self.ss_t = torch.empty(nss, B, dtype=F32, device=self.dev)
            self.ss_v = torch.empty(nss, B, dtype=F32, device=self.dev)
        with br("v"):   # image tower branch: projection + row norms
            if self.P is not None:
                vp_ = inp["v_pool"]
                gemm(vp_, vp_.stride(0), 0, w["W_v"], E, 0, Yv, self.P, 1, B, self.P, E, D_lo=Yvl,      # HF :262 visual_projection
                     row_ss=self.ss_v if fused_norm else None)
            if not fused_norm:
                it.norm_v(Yv, ldv, V_lo=Yvl)
        if self.P is not None:
            tp_ = inp["t_pool"]
            gemm(tp_, tp_.stride(0), 0, w["W_t"], E, 0, Yt, self.P, 1, B, self.P, E, D_lo=Ytl,          # HF :265 text_projection
                 row_ss=self.ss_t if fused_norm else None)
        if not fused_norm:
            it.norm_t(Yt, ldt, T_lo=Ytl)
        br.join("v")
        # small batch, fused step: forward AND gradient-operand tiles in one cluster launch (S stays in TMEM in between)
        self._itc_bwd_fused = bool(with_loss and getattr(self, "_in_step", False) and self.fuse_itc_small and it.can_fuse_small)
        if self._itc_bwd_fused:
            it.fwd_bwd_fused(Yt, ldt, Yv, ldv, self.scale, self.g_itc / (2.0 * B), T_lo=Ytl, V_lo=Yvl,
                             ss_t=self.ss_t if fused_norm else None, ss_v=self.ss_v if fused_norm else None)
        else:
            it.fwd_tiles(Yt, ldt, Yv, ldv, self.scale, T_lo=Ytl, V_lo=Yvl, ss_t=self.ss_t if fused_norm else None,
                         ss_v=self.ss_v if fused_norm else None)
        self._join_small_zero()        # itc_sums / r_sum live in the small zeroed block
        if with_loss:
            if it.can_inline_lse:      # small batch: the backward derives lse itself; the loss runs on a side branch
                with br("l"):
                    it.lse_loss(self.scale, self.z["itc_sums"])
            else:
                it.lse_loss(self.scale, self.z["itc_sums"])
        if it.logits is not None:
            self.out["logits_per_text"] = it.logits

    def _itc_bwd(self, inp, dS=None):
        """dS None: fused loss (tile recompute, g = beta_itc); else operands from the upstream gradient of the logits."""
        self._join_zero()
        B, E, w, z, o, it = self.B, self.E, self.w, self.z, self.out, self.itc
        Yt, Yv, Ytl, Yvl = self._itc_operands(inp)
        ldt, ldv = Yt.stride(0), Yv.stride(0)
        if dS is None:
            g = self.g_itc
            if not getattr(self, "_itc_bwd_fused", False):   # (else: emitted by the fused forward+backward launch already)
                it.bwd_operands(Yt, ldt, Yv, ldv, self.scale, g / (2.0 * B), T_lo=Ytl, V_lo=Yvl, inline_lse=it.can_inline_lse)
            dcoef = g / B
        else:
            assert it.precise, "autograd mode (materialised dS) is meant for drop-in batch sizes (< 4096)"
            call("tic_itc_ds_operands", ptr(dS), dS.stride(0), B, B, ptr(it.rinv_t), ptr(it.rinv_v), ptr(it.GA), ptr(it.GA_lo),
                 it.ld_ga, ptr(it.GBT), ptr(it.GBT_lo), it.ld_gbt, _stream())
            dcoef = 0.0
        br = self.br
        br.enabled = self.parallel_streams
        if self.P is not None:
            dYt, dYv, dYt_lo, dYv_lo = self.dY[:B], self.dY[B:], self.dY_lo[:B], self.dY_lo[B:]
            if not self.split:       # opt-in fast mode: single bf16 gradient operands (no residual K-segments)
                dYt_lo = dYv_lo = None
            tp_, vp_ = inp["t_pool"], inp["v_pool"]
            with br("v"):   # image-side gradient branch
                it.grad_gemm_v(Yt, ldt, T_lo=self._itc_gemm_lo(Ytl))
                it.finalize_v(it.acc_v, Yv, ldv, it.rinv_v, Yt, ldt, it.rinv_t, B, self.scale, dcoef, None, dYv,
                              dV_lo=dYv_lo, V_lo=Yvl, T_diag_lo=Ytl)
                gemm(dYv, self.P, 1, vp_, vp_.stride(0), 1, o["dW_v"], E, 0, self.P, E, B, A_lo=dYv_lo, accumulate=self._atomic["dW_v"])
            it.grad_gemm_t(Yv, ldv, V_lo=self._itc_gemm_lo(Yvl))
            it.finalize_t(Yt, ldt, Yv, ldv, it.rinv_v, self.scale, dcoef, None, dYt, z["r_sum"], dT_lo=dYt_lo, T_lo=Ytl,
                          V_diag_lo=Yvl)
            with br("w"):   # dW_t = dYt^T t_pool (both operands read MN-major)  ||  d_t_pool = dYt W_t
                gemm(dYt, self.P, 1, tp_, tp_.stride(0), 1, o["dW_t"], E, 0, self.P, E, B, A_lo=dYt_lo, accumulate=self._atomic["dW_t"])
            gemm(dYt, self.P, 0, w["W_t"], E, 1, o["d_t_pool"], E, 0, B, E, self.P, A_lo=dYt_lo)
            br.join("v")
            br.join("w")
        else:
            with br("v"):
                it.grad_gemm_v(Yt, ldt, T_lo=Ytl)
                it.finalize_v(it.acc_v, Yv, ldv, it.rinv_v, Yt, ldt, it.rinv_t, B, self.scale, dcoef, o["d_v_emb"], None)
            it.grad_gemm_t(Yv, ldv, V_lo=Yvl)
            it.finalize_t(Yt, ldt, Yv, ldv, it.rinv_v, self.scale, dcoef, o["d_t_emb"], None, z["r_sum"])
            br.join("v")

    def step(self, inp: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Fused forward + losses + backward.  inp (device tensors): t_pool, v_pool bf16 [B,E]; x_t bf16 [B,Lt,E] or
        [B,1,E]; x_v bf16 [B,Lv,E]; y_soft fp32 [B,C]; class_w fp32 [C] (optional); ITM: u_coin,u_pick fp32 [B] (+ ids, mask
        int64 [B,Lt] to gather) or precomputed lbl_tim int64 [B] + src_idx int32 [B]; keep uint8 [B,E] + keep_scale
        (dropout, optional).

        The ITC chain and the fusion/heads chain are independent until the loss mix, so they are issued on two streams
        (fork/join with events; captured as parallel branches of one CUDA graph) — the small-batch step is latency-bound."""
        B, z, o = self.B, self.z, self.out
        if self.parallel_streams and self.hi_priority_chains:
            # the two latency-critical chains run on high-priority streams, the parameter-gradient branches on ordinary ones:
            # when an SM frees up, the next kernel of a chain gets it before a split-K weight-gradient CTA does
            caller = torch.cuda.current_stream()
            if self._hi0 is None:
                self._hi0 = torch.cuda.Stream(device=self.dev, priority=-1)
            self._hi0.wait_stream(caller)
            self.br.pdl_chains = self.pdl_chains
            prev = capi.load().tic_set_pdl(1) if self.pdl_chains else None
            try:
                with torch.cuda.stream(self._hi0):
                    self._step_body(inp)
            finally:
                if prev is not None:
                    capi.load().tic_set_pdl(prev)
            caller.wait_stream(self._hi0)
            return o
        self._step_body(inp)
        return o

    def _step_body(self, inp):
        self._in_step = True
        try:
            return self._step_body_inner(inp)
        finally:
            self._in_step = False
            self._refreshed_all = False

    # ------------------------------------------------------------------ code warm-up (small-batch step)
    # Measured (scripts/fused_in_step.py, profiles/r02_fused_in_step.txt): inside the replayed step the fused ITC kernel takes
    # 24 us, alone 14 us — and 14 us again inside the step if the SAME kernel (outputs to scratch) ran once after the L2 flush.
    # What is cold is the kernel's CODE: ~1 us per KB of first-executed instructions when they come from HBM.  A latency-bound
    # step cannot hide that behind data-parallel work, so the step issues, on a side branch at its head, ONE tiny instance of
    # each kernel of the two critical chains (128-row problems on scratch buffers): by the time the real launches arrive, their
    # code is L2-resident.  It is an instruction prefetch spelled as launches; TIC_CODE_WARM=0 turns it off (A/B switch).
    def _build_warm(self):
        dev = self.dev
        e = lambda *sh, dt=F32: torch.zeros(*sh, dtype=dt, device=dev)  # noqa: E731
        W = {"A": e(128, 128, dt=BF16), "B": e(128, 128, dt=BF16), "Al": e(128, 128, dt=BF16), "Bl": e(128, 128, dt=BF16),
             "Df": e(128, 128), "Df2": e(128, 128), "Db": e(128, 128, dt=BF16), "Dl": e(128, 128, dt=BF16), "bias": e(128), "u": torch.full((128,), 0.75, device=dev),
             "lbl": torch.zeros(128, dtype=torch.int64, device=dev), "src": torch.zeros(128, dtype=torch.int32, device=dev),
             "ss": e(capi.load().tic_gemm_rowss_parts(64), 128), "y": e(8, self.C), "ws": e(16 * 8), "los": e(4), "lg": e(16, 8),
             "r": e(4)}
        W["A"].fill_(0.01); W["B"].fill_(0.02); W["ss"].fill_(1.0)
        W["it"] = ItcPlan(128, 128, 64, dev, precise=True)
        W["it"].scale_dev = self.scale_t
        # multi-GPU symmetric form: separate forward / gradient-operand launches on a row block without column statistics
        W["it_rb"] = ItcPlan(128, 128, 64, dev, precise=True, need_dv=False, col_sums=False)
        W["it_rb"].lse_row.fill_(3.0); W["it_rb"].lse_col.fill_(3.0)
        W["it_rb"].rinv_t.fill_(1.0); W["it_rb"].rinv_v.fill_(1.0)
        return W

    def _warm_calls(self, grp):
        """grp 0: the fused ITC tiles (the longest cold start, first needed);  1: the gradient-GEMM forms + normalise-backward;
        2: the fusion chain's heads / pairwise-gradient kernels"""
        if self._warm is None:
            self._warm = self._build_warm()
        W, E = self._warm, self.E
        A, Bm, Al, Bl, it = W["A"], W["B"], W["Al"], W["Bl"], W["it"]
        if grp == 0 and self.use_itc:
            if type(self).__name__ != "HeadPlan":      # multi-GPU plans: the un-fused tile kernels of the row / swapped blocks
                rb = W["it_rb"]
                rb.fwd_tiles(A, 128, Bm, 128, self.scale, T_lo=Al, V_lo=Bl)
                rb.bwd_operands(A, 128, Bm, 128, self.scale, 1e-3, T_lo=Al, V_lo=Bl)
            elif it.can_fuse_small and self.fuse_itc_small:
                it.fwd_bwd_fused(A, 128, Bm, 128, self.scale, 1e-3, T_lo=Al, V_lo=Bl, ss_t=W["ss"], ss_v=W["ss"])
        if grp == 1 and self.use_itc:
            gemm(A, 128, 0, Bm, 128, 1, W["Df"], 128, 0, 128, 64, 64, A_lo=Al, B_lo=Bl)          # dT / d_t_pool / dX form
            gemm(A, 128, 1, Bm, 128, 1, W["Df2"], 128, 0, 128, 64, 64, A_lo=Al, B_lo=Bl)         # dV / weight-gradient form
            call("tic_itc_grad_finalize", ptr(W["Df"]), 128, ptr(A), ptr(Al), 128, ptr(it.rinv_t), ptr(Bm), ptr(Bl), 128,
                 ptr(it.rinv_v), 8, 64, float(self.scale), 0.0, None, 128, ptr(W["Db"]), ptr(W["Dl"]), 128, ptr(W["r"]), 0,
                 ptr(self.scale_t), _stream())
        if grp == 2 and self.fusion in ("concat", "attention", "gmu"):
            if self.pairwise:
                call("tic_heads_fwd_bwd", ptr(self._warm_H), E, 1, E, self.C, int(self.use_itm), ptr(self.w["W_cls"]),
                     ptr(self.w["b_cls"]), ptr(self.w["W_tim"]), ptr(self.w["b_tim"]), ptr(W["y"]), None, ptr(W["lbl"]), None, 1.0,
                     0.0, 0.0, ptr(W["lg"]), ptr(W["lg"]), ptr(W["los"]), ptr(self._warm_dH), ptr(self._warm_dH_lo), E, None, E,
                     None, None, None, None, 1, ptr(W["ws"]), None, ptr(self._warm_P), ptr(self._warm_P), E, ptr(W["src"]), _stream())
                call("tic_fusion_pair_grad", ptr(self._warm_dH), ptr(self._warm_dH_lo), E, 1, E, int(self.use_itm), ptr(W["src"]),
                     ptr(self._warm_dP[0]), ptr(self._warm_dP[1]), ptr(self._warm_dP[2]), ptr(self._warm_dP[3]), E, _stream())

    def _step_body_inner(self, inp):
        B, z, o = self.B, self.z, self.out
        s0 = torch.cuda.current_stream()
        # Live weights: ONE refresh launch is the root of the step (bf16 working copies of both chains' matrices,
        # exp(logit_scale), the small accumulator block zeroed).  With one root per chain the graph launched the second
        # chain's root ~14 us late (CUPTI timeline, profiles/r02_timeline_c2_*.txt).
        self._refreshed_all = False
        if self._refresh_zeroes and "all" in self._refresh_groups:
            self._refresh("all")
            self._refreshed_all = True
        self._zero_accumulators()
        if self.code_warm and self.w and (self._refreshed_all or not self.live_weights):
            # three short side branches BEHIND the root — the refresh launch, or the accumulator memset of the snapshot /
            # multi-GPU plans (a captured step with several roots starts its later roots ~14 us late)
            self.br.enabled = True
            for grp in (0, 1, 2):
                with self.br("cw%d" % grp):
                    self._warm_calls(grp)
        two = self.use_itc and self.fusion is not None and self.parallel_streams
        if two:
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.dev, priority=-1 if self.hi_priority_chains else 0)
            s1 = self._side
            if self.itm_mode == 1:       # hard negatives read the materialised logits: sampling waits for the ITC tiles
                self._itc_fwd(inp)
            s1.wait_stream(s0)
            with torch.cuda.stream(s1):
                self._fusion_chain(inp)
            if self.itm_mode != 1:
                self._itc_fwd(inp)
            mixed = self._loss_mix_early()
            self._itc_bwd(inp)
            s0.wait_stream(s1)
        else:
            mixed = False
            if self.use_itc:
                self._itc_fwd(inp)
            if self.fusion is not None:
                self._fusion_chain(inp)
            elif self.use_itc:
                mixed = self._loss_mix_early()
            if self.use_itc:
                self._itc_bwd(inp)
        self.br.join("l")      # ITC loss terms (+ the early loss mix) on their side branch
        for grp in (0, 1, 2):
            self.br.join("cw%d" % grp)
        if not mixed:
            self._loss_mix()
        return o

    def _loss_mix(self):
        z, o = self.z, self.out
        call("tic_loss_mix", ptr(z["losses"]), ptr(z["itc_sums"]), getattr(self, "n_global", self.B),
             self.beta_itc if self.fusion is not None else 1.0, self.beta_itm, int(self.use_itc), int(self.use_itm),
             ptr(o["loss"]), _stream())

    def _loss_mix_early(self):
        """The mixed loss depends on the forward halves only (heads' loss sums, ITC lse sums): issue it on the `l` side branch
        as soon as both are there instead of after the backward of both chains, where it was one more serialised ~3 us
        launch at the tail of the step.  Base single-GPU plan only (the multi-GPU plans finish their lse sums elsewhere)."""
        base = self.itc.can_inline_lse and type(self)._itc_fwd is HeadPlan._itc_fwd
        # multi-GPU symmetric form: the rank's lse sums are complete when its forward returns (tic_itc_lse_rows)
        sym = getattr(self, "itc_mode", None) == "symmetric"
        if not (self._early_mix and self.parallel_streams and self.use_itc and (base or sym)):
            return False
        self.br.enabled = True
        with self.br("l"):          # same side stream as lse_loss: ordered after it
            if self.fusion is not None:
                torch.cuda.current_stream().wait_event(self._ev_heads)
            self._loss_mix()
        return True

    @property
    def _refresh_zeroes(self):
        """live weights: the refresh launches at the head of the two chains zero the small block (base single-GPU plan), or
        the ONE root launch zeroes all of it (any plan whose matrices fit one launch — the multi-GPU plans touch the block
        earlier in their chains, which only a root can precede)"""
        base = type(self)._itc_fwd is HeadPlan._itc_fwd
        return (self.live_weights and getattr(self, "_in_step", False) and (base or "all" in self._refresh_groups) and
                (not self.use_itc or "itc" in self._refresh_groups) and (self.fusion is None or "fusion" in self._refresh_groups))

    def _zero_accumulators(self):
        # The small block (loss sums, small-head gradients) is first touched by lse_loss / heads, several kernels into the
        # step: its memset leaves the head of both chains too (base plan only: the multi-GPU plans touch it earlier).
        self._zs_pending = False
        if self._refresh_zeroes:
            pass          # zeroed by tic_refresh_weights, the first kernel of each chain
        elif self.parallel_streams and self._defer_small_zero and type(self)._itc_fwd is HeadPlan._itc_fwd:
            self.br.enabled = True
            with self.br("zs"):
                self.zb_small.zero_()
            self._zs_pending = True
        else:
            self.zb_small.zero_()
        if not any(self._atomic.values()):
            self._z_pending = False                    # every large gradient is written by plain stores: nothing to zero
        elif self.parallel_streams and self.zb_big.numel() > 0:
            self.br.enabled = True
            with self.br("z"):
                self.zb_big.zero_()
            self._z_pending = True
        else:
            self.zb_big.zero_()
            self._z_pending = False

    def _join_zero(self):
        """call on a stream right before it first touches a large accumulator (weight gradients, d_xt_cls)"""
        if getattr(self, "_z_pending", False):
            self.br.join("z")

    def _join_small_zero(self):
        """call on a stream right before it first touches the small accumulator block (loss sums, small-head gradients)"""
        if getattr(self, "_zs_pending", False):
            self.br.join("zs")

    def _lo(self, t):
        """the bf16 residual twin of an intermediate, or None when split precision is off for this plan"""
        return t if self.split else None

    def _inline_rule(self, inp):
        """Uniform ITM decisions can be re-derived inside the pack kernel from the uniforms (same rule, bit-exact), so the
        sampler/gather kernel runs beside the fusion forward instead of before it."""
        return self.use_itm and self.itm_mode == 0 and "src_idx" not in inp and inp.get("u_coin") is not None

    def _fusion_chain(self, inp):
        self._refresh("fusion")
        if self.use_itm:
            if self._inline_rule(inp):
                self.br.enabled = self.parallel_streams
                with self.br("s"):
                    self._sample_itm(inp)
            else:
                self._sample_itm(inp)
        self._fusion_fwd(inp)
        self.br.join("s")      # lbl_tim (heads) / src_idx (unpack) come from the sampler
        self._join_small_zero()
        self._heads(inp, dH_f32=self.dHf if self.fusion == "aspect-att" else None)
        if self._early_mix:    # the loss mix needs the heads' loss sums only: it does not wait for the backward (see _loss_mix_early)
            self._ev_heads = torch.cuda.Event()
            self._ev_heads.record(torch.cuda.current_stream())
        self._fusion_bwd(inp)
        self.br.join("hw")
        self.br.join("g")

    def forward(self, inp: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Autograd mode, forward half: logits_per_text (materialised), mm_features, out_cls, out_tim.  ITM decisions come
        from the caller (`lbl_tim`, `src_idx`) exactly as the reference passes `tim_inputs`."""
        assert self.itc is None or self.itc.logits is not None, "autograd mode needs materialize_logits=True"
        self.generation += 1
        self.zb.zero_()
        if self.use_itc:
            self._itc_fwd(inp, with_loss=False)
        if self.fusion is not None:
            self._refresh("fusion")
            if self.use_itm:
                self._sample_itm(inp)
            self._fusion_fwd(inp)
            self._heads(inp, forward_only=True)
        return self.out

    def backward(self, inp, d_out_cls=None, d_logits=None, d_out_tim=None, generation=None) -> Dict[str, torch.Tensor]:
        """Autograd mode, backward half: upstream gradients of the three outputs -> gradients of inputs and parameters.
        The forward activations live in this plan's buffers: `generation` (the value of self.generation right after the
        matching forward()) makes a backward that follows ANOTHER forward on the same plan fail loudly instead of
        differentiating the wrong activations."""
        if generation is not None and generation != self.generation:
            raise RuntimeError("HeadPlan.backward: the plan ran another forward (generation %d) after the one being "
                               "differentiated (%d); its activations are gone.  Use one plan per in-flight forward."
                               % (self.generation, generation))
        B, C = self.B, self.C
        self.zb.zero_()
        if self.fusion is not None:
            dz = self.dz_ext
            dz.zero_()
            if d_out_cls is not None:
                dz[:B, :C].copy_(d_out_cls)
            if self.use_itm and d_out_tim is not None:
                dz[B:, :2].copy_(d_out_tim)
            self._heads(inp, dH_f32=self.dHf if self.fusion == "aspect-att" else None, dz_ext=dz)
            self._fusion_bwd(inp)
            self.br.join("hw")
        if self.use_itc:
            if d_logits is None:
                d_logits = torch.zeros(B, B, dtype=F32, device=self.dev)
            self._itc_bwd(inp, dS=d_logits.to(F32).contiguous())
        return self.out

    def _sample_itm(self, inp):
        B, o, st = self.B, self.out, _stream()
        if "src_idx" in inp:   # decisions made by the caller (e.g. the numpy-stream replay of the CLI)
            o["lbl_tim"].copy_(inp["lbl_tim"])
            o["src_idx"].copy_(inp["src_idx"].to(torch.int32))
            return
        ids, mask = inp.get("ids"), inp.get("mask")
        if ids is not None and "tim_ids" not in o:
            o["tim_ids"], o["tim_mask"] = torch.empty_like(ids), torch.empty_like(mask)
        if self.itm_mode == 1:
            # hard negatives, tile-stream form: the forward tiles left per-part weight sums; locate the part of every
            # mismatch row's target, then recompute the tiles and walk that part (csrc/itc.cu: ItcPickEpi)
            it = self.itc
            Yt, Yv, Ytl, Yvl = self._itc_operands(inp)
            it.hard_locate(inp["u_coin"], inp["u_pick"], o["lbl_tim"], o["src_idx"])
            it.hard_pick(Yt, Yt.stride(0), Yv, Yv.stride(0), self.scale, o["src_idx"], T_lo=Ytl, V_lo=Yvl)
            if ids is not None:   # the gathered ids / mask feed nothing downstream here (no second encoder pass): side branch
                self.br.enabled = self.parallel_streams
                with self.br("g"):
                    for s_, d_ in ((ids, o["tim_ids"]), (mask, o["tim_mask"])):
                        call("tic_gather_rows", ptr(s_), s_.stride(0) * s_.element_size(), ptr(d_),
                             d_.stride(0) * d_.element_size(), s_.shape[1] * s_.element_size(), ptr(o["src_idx"]), B, _stream())
            return
        if ids is not None:
            call("tic_itm_sample_gather", ptr(inp["u_coin"]), ptr(inp["u_pick"]), B, 0, None, 0, 0.0, None, ptr(ids),
                 ptr(mask), ids.stride(0) * ids.element_size(), ptr(o["tim_ids"]), ptr(o["tim_mask"]), ptr(o["lbl_tim"]),
                 ptr(o["src_idx"]), st)
        else:
            call("tic_itm_sample", ptr(inp["u_coin"]), ptr(inp["u_pick"]), B, 0, None, 0, 0.0, None, ptr(o["lbl_tim"]),
                 ptr(o["src_idx"]), st)

    def _heads(self, inp, dH_f32=None, forward_only=False, dz_ext=None):
        B, E, z, o, w = self.B, self.E, self.z, self.out, self.w
        y_soft = inp.get("y_soft")
        if y_soft is None:   # autograd mode: the classification loss is the caller's
            y_soft = self.y_dummy
        no = forward_only
        pw = getattr(self, "_pairwise", None) or (None, None, 0, None)   # (Pt, Pv, ldp, src): pairwise concat fusion
        side = self.parallel_streams and not no      # dW/db of the two small heads run beside the input-gradient chain
        wg = no or side
        call("tic_heads_fwd_bwd", ptr(self.H), E, B, E, self.C, int(self.use_itm), ptr(w["W_cls"]), ptr(w["b_cls"]),
             ptr(w["W_tim"]), ptr(w["b_tim"]), ptr(y_soft), ptr(inp.get("class_w")), ptr(o["lbl_tim"]),
             ptr(inp.get("keep")), float(inp.get("keep_scale", 1.0)), float(self.w_cls), float(self.beta_itm),
             ptr(o["out_cls"]), ptr(o.get("out_tim")), ptr(z["losses"]), None if no else ptr(self.dHb),
             None if no else ptr(self._lo(self.dHb_lo)), E, None if no else ptr(dH_f32), E, None if wg else ptr(z["dW_cls"]),
             None if wg else ptr(z["db_cls"]), None if wg else ptr(z["dW_tim"]), None if wg else ptr(z["db_tim"]), 1,
             ptr(self.heads_ws), ptr(dz_ext), ptr(pw[0]), ptr(pw[1]), pw[2], ptr(pw[3]), _stream())
        if side:
            self.br.enabled = True
            with self.br("hw"):
                call("tic_heads_wgrad", ptr(self.H), E, B, E, self.C, int(self.use_itm), ptr(self.heads_ws), ptr(inp.get("keep")),
                     float(inp.get("keep_scale", 1.0)), ptr(z["dW_cls"]), ptr(z["db_cls"]), ptr(z["dW_tim"]), ptr(z["db_tim"]),
                     _stream())

    def _fusion_fwd(self, inp):
        B, E, R, w, o, st = self.B, self.E, self.R, self.w, self.out, _stream()
        E2 = 2 * E
        src = o["src_idx"] if self.use_itm else None
        if self.fusion == "aspect-att":
            tp_, vp_ = inp["t_pool"], inp["v_pool"]
            call("tic_aspect_fwd", ptr(tp_), tp_.stride(0), ptr(vp_), vp_.stride(0), B, E, ptr(w["w_a"]), ptr(w["b_a"]),
                 ptr(self.H), E, ptr(self.alpha), st)
            return
        x_t, x_v = inp["x_t"], inp["x_v"]
        xt_stride, xv_stride = x_t.stride(0), x_v.stride(0)   # CLS rows: x[:,0,:]
        if self.pairwise:
            # linear_fusion([x_t | x_v]) = x_t W_f[:, :E]^T + (x_v W_f[:, E:]^T + b_f): both halves once per sample, on two
            # branches; the heads kernel forms relu(Pt[src] + Pv) for the main and the ITM rows itself (mm_late.py:92-96,170-181)
            self.br.enabled = self.parallel_streams
            with self.br("pv"):
                gemm(x_v, xv_stride, 0, w["W_f"].data_ptr() + 2 * E, E2, 0, self.Pv, E, 0, B, E, E, bias=w["b_f"])
            gemm(x_t, xt_stride, 0, w["W_f"], E2, 0, self.Pt, E, 0, B, E, E)
            self.br.join("pv")
            self._pairwise = (self.Pt, self.Pv, E, src)
            return
        self._pairwise = None
        X = self.Xcat
        X_lo = None
        uc, up = (ptr(inp["u_coin"]), ptr(inp["u_pick"])) if self._inline_rule(inp) else (None, None)
        if self.fusion == "concat" or self.fusion == "gmu":
            call("tic_pack_cls_pairs", ptr(x_t), xt_stride, ptr(x_v), xv_stride, B, E, ptr(src), ptr(X), E2, uc, up, st)
        if self.fusion == "attention":
            Ea, Lv = E + 8, self.Lv
            X_lo = self._lo(self.Xcat_lo)
            call("tic_pack_cls_pairs", ptr(x_t), xt_stride, None, 0, B, E, ptr(src), ptr(X), E2, uc, up, st)
            gemm(X, E2, 0, w["W_Q"], E, 0, self.q0, E, 1, R, E, E, bias=w["b_Q"], D_lo=self._lo(self.q0_lo))  # q0 = fc_Q(x_t[:,0])
            gemm(self.q0, E, 0, w["W_Kaug"], Ea, 1, self.kq, Ea, 0, R, Ea, E, A_lo=self._lo(self.q0_lo))       # [W_K^T q0 | <q0,b_K>] fp32
            call("tic_attn_pool_fwd", ptr(x_v), x_v.stride(0), x_v.stride(1), ptr(self.kq), Ea, B, 2 if self.use_itm else 1,
                 Lv, E, float(E) ** -0.5, ptr(self.xbar_b), ptr(self._lo(self.xbar_lo)), E, ptr(self.xbar_f), E, ptr(self.attn), Lv, st)
            gemm(self.xbar_b, E, 0, w["W_V"], E, 0, X.data_ptr() + 2 * E, E2, 1, R, E, E, bias=w["b_V"], A_lo=self._lo(self.xbar_lo),
                 D_lo=(X_lo.data_ptr() + 2 * E) if X_lo is not None else None)                                                      # ctx0 -> Xcat[:,E:]
        self.Hin, self.Hin_lo = X, X_lo
        if self.fusion == "gmu":
            gemm(X, E2, 0, w["W_gt"], E, 0, self.tp, E2, 0, R, E2, E, bias=w["b_gt"])               # linear_gmu_t(x_t[:,0])
            gemm(X.data_ptr() + 2 * E, E2, 0, w["W_gv"], E, 0, self.vp, E2, 0, R, E2, E, bias=w["b_gv"])
            call("tic_gmu_gate_fwd", ptr(X), E2, ptr(self.tp), ptr(self.vp), E2, R, E2, ptr(self.G), ptr(self._lo(self.G_lo)), E2, st)
            self.Hin, self.Hin_lo = self.G, self._lo(self.G_lo)
        gemm(self.Hin, E2, 0, w["W_f"], E2, 0, self.H, E, 0, R, E, E2, bias=w["b_f"], relu=True, A_lo=self.Hin_lo)

    def _fusion_bwd(self, inp):
        self._join_zero()      # the large accumulators (dW_*, d_xt_cls) are zero from here on
        B, E, R, w, z, o, st = self.B, self.E, self.R, self.w, self.z, self.out, _stream()
        E2 = 2 * E
        src = o["src_idx"] if self.use_itm else None
        if self.fusion == "aspect-att":
            tp_, vp_ = inp["t_pool"], inp["v_pool"]
            call("tic_aspect_bwd", ptr(tp_), tp_.stride(0), ptr(vp_), vp_.stride(0), B, E, ptr(w["w_a"]), ptr(w["b_a"]),
                 ptr(self.H), E, ptr(self.alpha), ptr(self.dHf), E, ptr(o["d_t_pool_fusion"]), E, ptr(z["dw_a"]),
                 ptr(z["db_a"]), st)
            return
        x_v = inp["x_v"]
        dH, dHl = self.dHb, self._lo(self.dHb_lo)
        lo = self._lo
        br = self.br
        br.enabled = self.parallel_streams
        if self.pairwise:
            x_t = inp["x_t"]
            with br("f"):     # bias gradient: column sums of dH over all R rows
                call("tic_colsum_bf16_pair", ptr(dH), ptr(dHl), E, R, E, ptr(z["db_f"]), _stream())
            call("tic_fusion_pair_grad", ptr(dH), ptr(dHl), E, B, E, int(self.use_itm), ptr(src), ptr(self.dPt), ptr(lo(self.dPt_lo)),
                 ptr(self.dPv), ptr(lo(self.dPv_lo)), E, st)
            with br("fv"):    # dW_f[:, E:] = dPv^T x_v_cls
                gemm(self.dPv, E, 1, x_v, x_v.stride(0), 1, o["dW_f"].data_ptr() + 4 * E, E2, 0, E, E, B, A_lo=lo(self.dPv_lo),
                     accumulate=self._atomic["dW_f"])
            with br("ft"):    # dW_f[:, :E] = dPt^T x_t_cls
                gemm(self.dPt, E, 1, x_t, x_t.stride(0), 1, o["dW_f"], E2, 0, E, E, B, A_lo=lo(self.dPt_lo),
                     accumulate=self._atomic["dW_f"])
            gemm(self.dPt, E, 0, w["W_f"], E2, 1, o["d_xt_cls"], E, 0, B, E, E, A_lo=lo(self.dPt_lo))   # d x_t[:,0] = dPt W_f[:, :E]
            br.join("f"); br.join("fv"); br.join("ft")
            return
        X, Hin, Hin_lo = self.Xcat, self.Hin, self.Hin_lo
        # backward through linear_fusion (dH is a split bf16 pair: hi + lo)
        with br("f"):   # parameter gradients of linear_fusion run beside the input-gradient chain
            call("tic_colsum_bf16_pair", ptr(dH), ptr(dHl), E, R, E, ptr(z["db_f"]), _stream())     # hi + lo in one launch
            gemm(dH, E, 1, Hin, E2, 1, o["dW_f"], E2, 0, E, E2, R, A_lo=dHl, B_lo=Hin_lo, accumulate=self._atomic["dW_f"])           # dW_f = dH^T Hin
        if self.fusion == "concat":
            gemm(dH, E, 0, w["W_f"], E2, 1, self.dXt, E, 0, R, E, E, A_lo=dHl)                      # dX_text = dH W_f[:, :E]
            call("tic_unpack_cls_grad", ptr(self.dXt), E, None, 0, B, E, ptr(src), ptr(o["d_xt_cls"]), E, st)
        elif self.fusion == "attention":
            Ea, Lv = E + 8, self.Lv
            gemm(dH, E, 0, w["W_f"], E2, 1, self.dXt, E, 0, R, E, E, A_lo=dHl)
            gemm(dH, E, 0, w["W_f"].data_ptr() + 2 * E, E2, 1, self.dctx, E, 1, R, E, E, A_lo=dHl, D_lo=lo(self.dctx_lo))
            call("tic_colsum_bf16_pair", ptr(self.dctx), ptr(lo(self.dctx_lo)), E, R, E, ptr(z["db_V"]), st)
            gemm(self.dctx, E, 1, self.xbar_b, E, 1, o["dW_V"], E, 0, E, E, R, A_lo=lo(self.dctx_lo), B_lo=lo(self.xbar_lo), accumulate=self._atomic["dW_V"])
            gemm(self.dctx, E, 0, w["W_V"], E, 1, self.dxbar, E, 0, R, E, E, A_lo=lo(self.dctx_lo))
            call("tic_attn_pool_bwd", ptr(x_v), x_v.stride(0), x_v.stride(1), ptr(self.attn), Lv, ptr(self.dxbar), E,
                 ptr(self.xbar_f), E, B, 2 if self.use_itm else 1, Lv, E, float(E) ** -0.5, ptr(self.dkq), ptr(lo(self.dkq_lo)),
                 Ea, st)
            gemm(self.dkq, Ea, 0, w["W_Kaug"], Ea, 0, self.dq0, E, 1, R, E, Ea, A_lo=lo(self.dkq_lo), D_lo=lo(self.dq0_lo))
            gemm(self.q0, E, 1, self.dkq, Ea, 1, self.dWK_aug, Ea, 0, E, Ea, R, A_lo=lo(self.q0_lo), B_lo=lo(self.dkq_lo), accumulate=self._atomic["dWK_aug"])  # [dW_K|db_K]
            call("tic_colsum_bf16_pair", ptr(self.dq0), ptr(lo(self.dq0_lo)), E, R, E, ptr(z["db_Q"]), st)
            gemm(self.dq0, E, 1, X, E2, 1, o["dW_Q"], E, 0, E, E, R, A_lo=lo(self.dq0_lo), accumulate=self._atomic["dW_Q"])
            gemm(self.dq0, E, 0, w["W_Q"], E, 1, self.dXt2, E, 0, R, E, E, A_lo=lo(self.dq0_lo))
            call("tic_unpack_cls_grad", ptr(self.dXt), E, ptr(self.dXt2), E, B, E, ptr(src), ptr(o["d_xt_cls"]), E, st)
        elif self.fusion == "gmu":
            gemm(dH, E, 0, w["W_f"], E2, 1, self.dG, E2, 0, R, E2, E, A_lo=dHl)                     # dG = dH W_f
            call("tic_gmu_gate_bwd", ptr(X), E2, ptr(self.tp), ptr(self.vp), E2, ptr(self.dG), E2, R, E2, ptr(self.dtp),
                 ptr(self.dvp), ptr(lo(self.dtp_lo)), ptr(lo(self.dvp_lo)), E2, ptr(self.dXg), E2, st)
            call("tic_colsum_bf16_pair", ptr(self.dtp), ptr(lo(self.dtp_lo)), E2, R, E2, ptr(z["db_gt"]), st)
            call("tic_colsum_bf16_pair", ptr(self.dvp), ptr(lo(self.dvp_lo)), E2, R, E2, ptr(z["db_gv"]), st)
            gemm(self.dtp, E2, 1, X, E2, 1, o["dW_gt"], E, 0, E2, E, R, A_lo=lo(self.dtp_lo), accumulate=self._atomic["dW_gt"])
            gemm(self.dvp, E2, 1, X.data_ptr() + 2 * E, E2, 1, o["dW_gv"], E, 0, E2, E, R, A_lo=lo(self.dvp_lo), accumulate=self._atomic["dW_gv"])
            gemm(self.dtp, E2, 0, w["W_gt"], E, 1, self.dXt2, E, 0, R, E, E2, A_lo=lo(self.dtp_lo))
            call("tic_unpack_cls_grad", ptr(self.dXg), E2, ptr(self.dXt2), E, B, E, ptr(src), ptr(o["d_xt_cls"]), E, st)
        br.join("f")

class HostStep:
    """Host-facing entry point: the step's inputs live in HOST memory (what a reference-side caller holds).  All inputs
    share one pinned arena (`host_views[name]` are typed views the caller may fill in place) mirrored by one device arena,
    so a call is: [optional copy of caller tensors into the arena] -> ONE host->device copy of `h2d_bytes` -> the plan
    (replayed as a CUDA graph) -> device->host copy of the 4 losses (`d2h_bytes`) -> stream synchronise."""

    def __init__(self, plan, host_example: Dict[str, torch.Tensor], bf16_keys=(), use_graph: bool = True):
        self.plan, self.bf16_keys = plan, tuple(bf16_keys)
        dev = plan.dev
        layout, off = {}, 0
        for k, v in host_example.items():
            dt = BF16 if k in self.bf16_keys else v.dtype
            nbytes = v.numel() * torch.empty(0, dtype=dt).element_size()
            layout[k] = (off, nbytes, dt, tuple(v.shape))
            off = (off + nbytes + 255) // 256 * 256
        self.h2d_bytes = sum(n for _, n, _, _ in layout.values())
        self.arena_bytes = off
        self.pinned = torch.empty(off, dtype=torch.uint8, pin_memory=True)
        self.device = torch.empty(off, dtype=torch.uint8, device=dev)
        self.host_views = {k: self.pinned[o:o + n].view(dt).view(shape) for k, (o, n, dt, shape) in layout.items()}
        self.dev_in = {k: self.device[o:o + n].view(dt).view(shape) for k, (o, n, dt, shape) in layout.items()}
        for k, v in host_example.items():
            self.host_views[k].copy_(v.to(self.host_views[k].dtype))
        self.loss_pinned = torch.empty(4, dtype=F32, pin_memory=True)
        self.d2h_bytes = self.loss_pinned.numel() * 4
        self.graph = None
        if use_graph:
            self.device.copy_(self.pinned)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                plan.step(self.dev_in)
            torch.cuda.current_stream().wait_stream(s)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                plan.step(self.dev_in)

    def __call__(self, host: Optional[Dict[str, torch.Tensor]] = None):
        if host is not None:
            for k, v in host.items():
                hv = self.host_views[k]
                hv.copy_(v if v.dtype == hv.dtype else v.to(hv.dtype))
        self.device.copy_(self.pinned, non_blocking=True)        # the step's inputs: one H2D transfer
        if self.graph is not None:
            self.graph.replay()
        else:
            self.plan.step(self.dev_in)
        self.loss_pinned.copy_(self.plan.out["loss"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return [float(x) for x in self.loss_pinned]


class HostPipeline:
    """Pipelined host-facing entry point: `submit()` enqueues one step whose inputs live in a pinned HOST arena (H2D copy
    on a copy stream -> the plan replayed as a CUDA graph on the compute stream -> D2H of the 4 losses) and returns at once;
    `result()` blocks until the oldest outstanding step has finished and returns its losses.  Two slots (host arena, device
    arena, graph) alternate, so the H2D transfer of step k+1 overlaps the kernels of step k — what a training loop whose
    encoders run elsewhere would do.  Every step still pays its own H2D and D2H inside the pipeline.
    `between` (optional callable) is enqueued on the compute stream before each step (bench.py: the L2 flush)."""

    DEPTH = 2

    def __init__(self, plan, host_example: Dict[str, torch.Tensor], bf16_keys=(), between=None):
        self.plan, self.between = plan, between
        self.slots = [HostStep(plan, host_example, bf16_keys=bf16_keys, use_graph=True) for _ in range(self.DEPTH)]
        self.h2d_bytes, self.d2h_bytes = self.slots[0].h2d_bytes, self.slots[0].d2h_bytes
        self.copy_stream = torch.cuda.Stream(device=plan.dev)
        self.h2d_done = [torch.cuda.Event() for _ in self.slots]
        self.step_done = [torch.cuda.Event() for _ in self.slots]
        self.n_submitted = self.n_collected = 0

    def host_views(self, k=None):
        """Pinned typed views the caller fills in place for the NEXT submitted step."""
        return self.slots[(self.n_submitted if k is None else k) % self.DEPTH].host_views

    def submit(self):
        assert self.n_submitted - self.n_collected < self.DEPTH, "collect a result before submitting another step"
        i = self.n_submitted % self.DEPTH
        sl, cs = self.slots[i], torch.cuda.current_stream()
        self.copy_stream.wait_event(self.step_done[i])          # the slot's previous step has consumed its device arena
        with torch.cuda.stream(self.copy_stream):
            sl.device.copy_(sl.pinned, non_blocking=True)       # this step's inputs: one H2D transfer
            self.h2d_done[i].record()
        if self.between is not None:
            self.between()
        cs.wait_event(self.h2d_done[i])
        sl.graph.replay()
        sl.loss_pinned.copy_(self.plan.out["loss"], non_blocking=True)
        self.step_done[i].record()
        self.n_submitted += 1

    def result(self):
        assert self.n_collected < self.n_submitted
        i = self.n_collected % self.DEPTH
        self.step_done[i].synchronize()
        self.n_collected += 1
        return [float(x) for x in self.slots[i].loss_pinned]
