"""Multi-GPU execution of the path (not present in the reference, which is single-device).

One process per GPU, torch.distributed (NCCL over NVLink/NVSwitch).  Only ITC has an exchange step:

    rank r owns text rows and image rows [r*b, (r+1)*b) of the global batch N = b * world
    1. all_gather  V (image embeddings, + bf16 residual and inverse norms)               [N, P]
    2. local       row block of similarity tiles  S[rows_r, :]  -> row sums (complete), column partial sums [N]
    3. all_reduce  column partial sums (SUM — the fixed softmax shift makes them plain sums)  [N] fp32
    4. local       lse, loss terms of my rows; backward operands GA [b, N], GBT [N, b]
    5. local       dT_acc = GA * V_all  [b, P];   dV_acc_all = GBT * T_r  [N, P]
    6. reduce_scatter(SUM) dV_acc_all -> my [b, P] block; finalise dT, dV locally
    7. the caller all-reduces weight gradients / d logit_scale (SUM): every rank produces gradients of the GLOBAL loss
       restricted to its samples.

Fusion heads, the uniform ITM sampler, gathers and the small heads are per-sample: pure data parallel, no collective.
`ShardedItc` holds the collective sequencing and is backend-agnostic: on GPUs the block backend is `plan.ItcPlan`
(CUDA kernels); the gloo/CPU tests plug in a torch stand-in to check the decomposition against the single-process oracle.
"""
from typing import Optional

import torch
import torch.distributed as dist


class ShardedItc:
    """Collective sequencing of the row-sharded ITC step. `block` implements the ItcPlan piece interface."""

    def __init__(self, block, b_local: int, world: int, rank: int, P: int, group=None):
        self.block, self.b, self.world, self.rank, self.P, self.group = block, b_local, world, rank, P, group
        self.N = b_local * world
        self.row_offset = rank * b_local

    def _all_gather(self, x: torch.Tensor) -> torch.Tensor:
        out = torch.empty((self.world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous(), group=self.group)
        return out

    def _reduce_scatter(self, full: torch.Tensor) -> torch.Tensor:
        mine = torch.empty((self.b,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
        if dist.get_backend(self.group) == "gloo":   # gloo has no reduce_scatter: all_reduce + slice (CPU tests only)
            tmp = full.clone()
            dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=self.group)
            mine.copy_(tmp[self.row_offset:self.row_offset + self.b])
        else:
            dist.reduce_scatter_tensor(mine, full.contiguous(), op=dist.ReduceOp.SUM, group=self.group)
        return mine

    def forward(self, T, V, scale, loss_sums, T_lo=None, V_lo=None):
        """T, V: this rank's [b, P] embeddings. Fills block.lse_*, adds this rank's loss terms to loss_sums[2]."""
        blk = self.block
        V_all = self._all_gather(V)                                                    # (1)
        V_all_lo = self._all_gather(V_lo) if V_lo is not None else None
        self.V_all, self.V_all_lo = V_all, V_all_lo
        ldt, ldv = T.stride(0), V_all.stride(0)
        blk.norms(T, ldt, V_all, ldv, T_lo=T_lo, V_lo=V_all_lo)                        # row norms are cheap: recomputed, not gathered
        blk.fwd_tiles(T, ldt, V_all, ldv, scale, T_lo=T_lo, V_lo=V_all_lo)             # (2)
        col_sum = blk.reduce_col_parts()                                               # [N] fp32, local partial
        dist.all_reduce(col_sum, op=dist.ReduceOp.SUM, group=self.group)               # (3)
        blk.lse_loss(scale, loss_sums, col_parts=col_sum, n_col_parts=1)               # (4)

    def backward(self, T, V, scale, g, dT_f32=None, dT_bf16=None, dV_f32=None, dV_bf16=None, r_sum=None, T_lo=None,
                 V_lo=None, dT_lo=None, dV_lo=None):
        """g = dLoss/d(clip_loss) of the GLOBAL loss.  Produces this rank's dT, dV [b, P]."""
        blk, N, b = self.block, self.N, self.b
        V_all, V_all_lo = self.V_all, self.V_all_lo
        ldt, ldv = T.stride(0), V_all.stride(0)
        blk.bwd_operands(T, ldt, V_all, ldv, scale, g / (2.0 * N), T_lo=T_lo, V_lo=V_all_lo)   # (4)
        blk.grad_gemms(T, ldt, V_all, ldv, T_lo=T_lo, V_lo=V_all_lo)                           # (5)
        acc_v_mine = self._reduce_scatter(blk.acc_v)                                           # (6)
        V_mine = V_all[self.row_offset:self.row_offset + b]
        V_mine_lo = V_all_lo[self.row_offset:self.row_offset + b] if V_all_lo is not None else None
        rinv_v_mine = blk.rinv_v[self.row_offset:self.row_offset + b]
        blk.finalize_t(T, ldt, V_mine, ldv, rinv_v_mine, scale, g / N, dT_f32, dT_bf16, r_sum, dT_lo=dT_lo, T_lo=T_lo,
                       V_diag_lo=V_mine_lo)
        blk.finalize_v(acc_v_mine, V_mine, ldv, rinv_v_mine, T, ldt, blk.rinv_t, b, scale, g / N, dV_f32, dV_bf16,
                       dV_lo=dV_lo, V_lo=V_mine_lo, T_diag_lo=T_lo)


def _make_dist_head_plan():
    from . import plan as P
    from .capi import call, ptr

    class _ItcBlock(P.ItcPlan):
        def reduce_col_parts(self):
            call("tic_reduce_parts", ptr(self.col_part), self.ncp, self.n, ptr(self.col_sum), P._stream())
            return self.col_sum

    class DistHeadPlan(P.HeadPlan):
        """HeadPlan whose ITC part is sharded over the process group; every rank computes gradients of the global loss
        restricted to its samples (sum the weight gradients across ranks to get the global gradient)."""

        def __init__(self, B_local, *, world, rank, group=None, d: Optional[int] = None, **kw):
            kw.setdefault("use_itc", True)
            use_itc = kw["use_itc"]
            kw["use_itc"] = False                      # the parent must not allocate a square single-GPU ItcPlan
            super().__init__(B_local, **kw)
            self.use_itc = use_itc
            self.world, self.rank = world, rank
            if self.P is None and d is not None:
                self.Pe = d
                self.out["d_t_emb"] = torch.empty(B_local, d, device=self.dev)
                self.out["d_v_emb"] = torch.empty(B_local, d, device=self.dev)
            N = B_local * world
            if use_itc:
                self.itc = _ItcBlock(B_local, N, self.Pe, self.dev, row_offset=rank * B_local)
                self.sharded = ShardedItc(self.itc, B_local, world, rank, self.Pe, group)
                self.beta_itc = kw.get("beta_itc", 0.1) if self.fusion is not None else 1.0
                self.g_itc = self.beta_itc if self.fusion is not None else 1.0
                self.w_cls = (1.0 - (self.beta_itc + self.beta_itm)) if self.fusion is not None else 0.0
            # local means -> contributions to global means
            self.w_cls /= world
            self.beta_itm_local = self.beta_itm / world
            self.n_global = N

        def _heads(self, inp, dH_f32=None, forward_only=False, dz_ext=None):
            saved = self.beta_itm
            self.beta_itm = self.beta_itm_local
            try:
                super()._heads(inp, dH_f32=dH_f32, forward_only=forward_only, dz_ext=dz_ext)
            finally:
                self.beta_itm = saved

        def _itc_fwd(self, inp, with_loss=True):
            B, E, w = self.B, self.E, self.w
            Yt, Yv, Ytl, Yvl = self._itc_operands(inp)
            if self.P is not None:
                tp_, vp_ = inp["t_pool"], inp["v_pool"]
                P.gemm(tp_, tp_.stride(0), 0, w["W_t"], E, 0, Yt, self.P, 1, B, self.P, E, D_lo=Ytl)
                P.gemm(vp_, vp_.stride(0), 0, w["W_v"], E, 0, Yv, self.P, 1, B, self.P, E, D_lo=Yvl)
            self.sharded.forward(Yt, Yv, self.scale, self.z["itc_sums"], T_lo=Ytl, V_lo=Yvl)

        def _itc_bwd(self, inp, dS=None):
            assert dS is None, "the sharded path implements the fused loss only"
            self._join_zero()
            B, E, w, z, o = self.B, self.E, self.w, self.z, self.out
            Yt, Yv, Ytl, Yvl = self._itc_operands(inp)
            if self.P is not None:
                dYt, dYv, dYt_lo, dYv_lo = self.dY[:B], self.dY[B:], self.dY_lo[:B], self.dY_lo[B:]
                self.sharded.backward(Yt, Yv, self.scale, self.g_itc, dT_bf16=dYt, dV_bf16=dYv, r_sum=z["r_sum"], T_lo=Ytl,
                                      V_lo=Yvl, dT_lo=dYt_lo, dV_lo=dYv_lo)
                tp_, vp_ = inp["t_pool"], inp["v_pool"]
                P.gemm(dYt, self.P, 1, tp_, tp_.stride(0), 1, o["dW_t"], E, 0, self.P, E, B, A_lo=dYt_lo, accumulate=self._atomic["dW_t"])
                P.gemm(dYv, self.P, 1, vp_, vp_.stride(0), 1, o["dW_v"], E, 0, self.P, E, B, A_lo=dYv_lo, accumulate=self._atomic["dW_v"])
                P.gemm(dYt, self.P, 0, w["W_t"], E, 1, o["d_t_pool"], E, 0, B, E, self.P, A_lo=dYt_lo)
            else:
                self.sharded.backward(Yt, Yv, self.scale, self.g_itc, dT_f32=o["d_t_emb"], dV_f32=o["d_v_emb"],
                                      r_sum=z["r_sum"])

        def step(self, inp):
            o = super().step(inp)
            return o

        def global_loss(self):
            """[mix, cls, itc, itm] of the GLOBAL batch (reporting only): cls/itm are means of the rank means, itc is the
            sum of the per-rank row contributions."""
            t = self.out["loss"].clone()
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.sharded.group if self.use_itc else None)
            cls, itc, itm = t[1] / self.world, t[2], t[3] / self.world
            if self.fusion is None:
                return torch.stack([itc, cls, itc, itm])
            mix = (1.0 - (self.beta_itc + self.beta_itm)) * cls + self.beta_itc * itc + self.beta_itm * itm
            return torch.stack([mix, cls, itc, itm])

    return DistHeadPlan


def __getattr__(name):
    if name == "DistHeadPlan":
        return _make_dist_head_plan()
    raise AttributeError(name)
