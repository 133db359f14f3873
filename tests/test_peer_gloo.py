"""CPU, world_size 2 over gloo: the SYMMETRIC row/column-block formulation of the sharded ITC step
(tic_b200.peer.SymmetricItc — two exchanges, no reduction across ranks) reproduces the single-process oracle on the
concatenated batch: loss, this rank's dT and dV, and the summed d logit_scale.  Same for the row-block form with a
peer reduction of the image-side gradient (tic_b200.peer.RowBlockItc, used from 4096 global negatives on).  The exchange kernel and the tile kernels
need GPUs, so the exchange here is a gloo all_gather and the per-rank block backend a torch (fp64) stand-in with the
piece interface of plan.ItcPlan; tests/dist_gpu_check.py --mode peer covers the real thing on 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restatement as R


class CpuBlock:
    """fp64 stand-in for one ItcPlan block: A rows are mine (m), B rows are gathered (n)."""

    def __init__(self, m, n, row_offset):
        self.m, self.n, self.row_offset = m, n, row_offset
        self.rinv_t = torch.zeros(m, dtype=torch.float64)     # published by norm_t
        self.rinv_v = torch.zeros(n, dtype=torch.float64)     # filled by exchange("emb")
        self.lse_row = torch.zeros(m, dtype=torch.float64)    # published by lse_rows
        self.lse_col = torch.zeros(n, dtype=torch.float64)    # filled by exchange("lse")

    def norm_t(self, T, ldt, T_lo=None):
        self.rinv_t.copy_(1.0 / T.norm(dim=1))

    def fwd_tiles(self, T, ldt, V, ldv, scale, T_lo=None, V_lo=None, seg=None):
        self.seen_seg_fwd = seg
        self.S = scale * (T * self.rinv_t[:, None]) @ (V * self.rinv_v[:, None]).t()
        self.row_sum = torch.exp(self.S - scale).sum(1)
        self.diag = self.S[torch.arange(self.m), self.row_offset + torch.arange(self.m)]

    def bwd_operands(self, T, ldt, V, ldv, scale, gscale, T_lo=None, V_lo=None, seg=None):
        self.seen_seg_bwd = seg
        Gp = gscale * (torch.exp(self.S - self.lse_row[:, None]) + torch.exp(self.S - self.lse_col[None, :]))
        self.GA = Gp * self.rinv_v[None, :]

    def grad_gemm_t(self, V, ldv, V_lo=None):
        self.acc_t = self.GA @ V

    def finalize_t(self, T, ldt, V_diag, ldv, rinv_v_diag, scale, diag_coef, dT_f32, dT_bf16, r_sum, **kw):
        dxh = scale * (self.acc_t - diag_coef * rinv_v_diag[:, None] * V_diag)
        xh = T * self.rinv_t[:, None]
        r = (xh * dxh).sum(1)
        dT_f32.copy_(self.rinv_t[:, None] * (dxh - xh * r[:, None]))
        if r_sum is not None:
            r_sum += r.sum()


def _worker(rank, world, port, b, P, out_q, push=None):
    """push: None = the pull form (exchange phases); "beside" / "before" = the PUSH form's sequencing (tic_peer_push): the
    embeddings are delivered by the "emb" callable (on a side branch / stream-ordered before the polling tile kernels), the
    lse vectors by the lse_rows launch itself, the tile kernels receive the flag tuples, and "done" is signalled once."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tic_b200.peer import SymmetricItc
        N = b * world
        g = torch.Generator().manual_seed(7)
        T_full = torch.randn(N, P, generator=g, dtype=torch.float64)
        V_full = torch.randn(N, P, generator=g, dtype=torch.float64) + 0.5 * T_full
        T, V = T_full[rank * b:(rank + 1) * b].clone(), V_full[rank * b:(rank + 1) * b].clone()
        T_all, V_all = torch.zeros(N, P, dtype=torch.float64), torch.zeros(N, P, dtype=torch.float64)
        scale = float(np.exp(2.6592))
        rb, cb = CpuBlock(b, N, rank * b), CpuBlock(b, N, rank * b)

        def exchange(phase):   # stand-in for tic_peer_exchange: gathers what every rank published
            if phase == "emb":
                pairs = [(T, T_all), (V, V_all), (rb.rinv_t, cb.rinv_v), (cb.rinv_t, rb.rinv_v)]
            else:
                pairs = [(rb.lse_row, cb.lse_col), (cb.lse_row, rb.lse_col)]
            for mine, full in pairs:
                dist.all_gather_into_tensor(full, mine.contiguous())

        def lse_rows(rb_, cb_, s, loss_sums):
            rb_.lse_row.copy_(s + torch.log(rb_.row_sum))
            cb_.lse_row.copy_(s + torch.log(cb_.row_sum))
            loss_sums[0] += (rb_.lse_row - rb_.diag).sum()
            loss_sums[1] += (cb_.lse_row - rb_.diag).sum()

        calls = []
        pu = None
        if push is not None:
            def lse_rows_push(rb_, cb_, s, loss_sums):      # the lse exchange rides on this launch in the push form
                lse_rows(rb_, cb_, s, loss_sums)
                exchange("lse")
                calls.append("lse_rows+push")
            pu = {"emb": lambda: (exchange("emb"), calls.append("emb")), "lse": lambda: calls.append("lse(no-op)"),
                  "done": lambda: calls.append("done"), "seg_emb": ("flags-emb", rank), "seg_lse": ("flags-lse", rank)}

            def no_pull(phase):
                raise AssertionError("the push form must not run the pull-form exchange %r" % phase)
            sym = SymmetricItc(rb, cb, no_pull, lse_rows_push, b, world, rank, push=pu)
            assert sym.push_beside            # tiny grids: the push runs beside the polling kernels
            sym.push_beside = push == "beside"
        else:
            sym = SymmetricItc(rb, cb, exchange, lse_rows, b, world, rank)
        sums = torch.zeros(2, dtype=torch.float64)
        sym.forward(T, V, T_all, V_all, scale, sums)
        dT, dV = torch.empty(b, P, dtype=torch.float64), torch.empty(b, P, dtype=torch.float64)
        rsum = torch.zeros(1, dtype=torch.float64)
        sym.backward(T, V, T_all, V_all, scale, 1.0, dT_f32=dT, dV_f32=dV, r_sum=rsum)
        dist.all_reduce(sums)
        dist.all_reduce(rsum)
        loss = 0.5 * (sums[0] + sums[1]) / N
        Tq, Vq = T_full.clone().requires_grad_(True), V_full.clone().requires_grad_(True)
        ls = torch.tensor(2.6592, dtype=torch.float64, requires_grad=True)
        ref = R.clip_loss(R.itc_logits(Tq, Vq, ls))
        ref.backward()
        ok = (abs(float(loss) - float(ref)) < 1e-10
              and torch.allclose(dT, Tq.grad[rank * b:(rank + 1) * b], rtol=1e-8, atol=1e-12)
              and torch.allclose(dV, Vq.grad[rank * b:(rank + 1) * b], rtol=1e-8, atol=1e-12)
              and abs(float(rsum) - float(ls.grad)) < 1e-9
              and torch.allclose(rb.diag, cb.diag, rtol=1e-12))
        if push is not None:     # sequencing of the push form
            ok = (ok and calls == ["emb", "lse_rows+push", "lse(no-op)", "done"]
                  and rb.seen_seg_fwd == pu["seg_emb"] and cb.seen_seg_fwd == pu["seg_emb"]
                  and rb.seen_seg_bwd == pu["seg_lse"] and cb.seen_seg_bwd == pu["seg_lse"])
        out_q.put((rank, bool(ok), float(loss), float(ref)))
    finally:
        dist.destroy_process_group()


class CpuRowBlock:
    """fp64 stand-in for the row-block backend of RowBlockItc (GA-shared mode: dV contributions carry rinv_v[j])."""

    def __init__(self, b, N, P, rank, world):
        self.b, self.N, self.rank, self.world, self.row_offset = b, N, rank, world, rank * b
        self.rinv_v = torch.zeros(N, dtype=torch.float64)
        self.rinv_pub = torch.zeros(b, dtype=torch.float64)
        self.col_sum = torch.zeros(N, dtype=torch.float64)
        self.col_all = torch.zeros(world, N, dtype=torch.float64)
        self.acc_v = torch.zeros(N, P, dtype=torch.float64)
        self.dv_parts = torch.zeros(world, b, P, dtype=torch.float64)

    def rinv_v_mine(self):
        return self.rinv_pub

    def publish_v_norm(self, V):
        self.rinv_pub.copy_(1.0 / V.norm(dim=1))

    def norm_t(self, T, ldt):
        self.rinv_t = 1.0 / T.norm(dim=1)
        self.That = T * self.rinv_t[:, None]

    def fwd_tiles(self, T, ldt, V, ldv, scale):
        self.S = scale * self.That @ (V * self.rinv_v[:, None]).t()
        E = torch.exp(self.S - scale)
        self.row_sum, self.col_part = E.sum(1), E.sum(0)
        self.diag = self.S[torch.arange(self.b), self.row_offset + torch.arange(self.b)]

    def publish_col_sums(self):
        self.col_sum.copy_(self.col_part)

    def lse_loss_gathered(self, scale, loss_sums):
        self.lse_row = scale + torch.log(self.row_sum)
        self.lse_col = scale + torch.log(self.col_all.sum(0))
        loss_sums[0] += (self.lse_row - self.diag).sum()
        loss_sums[1] += (self.lse_col[self.row_offset:self.row_offset + self.b] - self.diag).sum()

    def bwd_operands(self, T, ldt, V, ldv, scale, gscale):
        Gp = gscale * (torch.exp(self.S - self.lse_row[:, None]) + torch.exp(self.S - self.lse_col[None, :]))
        self.GA = Gp * self.rinv_v[None, :]

    def grad_gemm_v(self, T, ldt):
        self.acc_v.copy_(self.GA.t() @ self.That)

    def reduce_dv(self):
        return self.dv_parts.sum(0)

    def grad_gemm_t(self, V, ldv):
        self.acc_t = self.GA @ V

    @staticmethod
    def _fin(acc, X, rinv, Xo, rinv_o, scale, diag_coef):
        dxh = scale * (acc - diag_coef * rinv_o[:, None] * Xo)
        xh = X * rinv[:, None]
        r = (xh * dxh).sum(1)
        return rinv[:, None] * (dxh - xh * r[:, None]), r

    def finalize_v(self, acc_v, V, ldv, rinv_v, T_diag, ldt, rinv_t_diag, rows, scale, diag_coef, dV_f32, dV_bf16, **kw):
        g, _ = self._fin(acc_v / rinv_v[:, None], V, rinv_v, T_diag, rinv_t_diag, scale, diag_coef)
        dV_f32.copy_(g)

    def finalize_t(self, T, ldt, V_diag, ldv, rinv_v_diag, scale, diag_coef, dT_f32, dT_bf16, r_sum, **kw):
        g, r = self._fin(self.acc_t, T, self.rinv_t, V_diag, rinv_v_diag, scale, diag_coef)
        dT_f32.copy_(g)
        if r_sum is not None:
            r_sum += r.sum()


def _worker_rowblock(rank, world, port, b, P, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tic_b200.peer import RowBlockItc
        N = b * world
        g = torch.Generator().manual_seed(9)
        T_full = torch.randn(N, P, generator=g, dtype=torch.float64)
        V_full = torch.randn(N, P, generator=g, dtype=torch.float64) + 0.5 * T_full
        T, V = T_full[rank * b:(rank + 1) * b].clone(), V_full[rank * b:(rank + 1) * b].clone()
        V_all = torch.zeros(N, P, dtype=torch.float64)
        scale = float(np.exp(2.6592))
        blk = CpuRowBlock(b, N, P, rank, world)

        def exchange(phase):   # stand-in for tic_peer_exchange
            if phase == "emb":
                dist.all_gather_into_tensor(V_all, V.contiguous())
                dist.all_gather_into_tensor(blk.rinv_v, blk.rinv_pub)
            elif phase == "col":
                dist.all_gather_into_tensor(blk.col_all.view(-1), blk.col_sum)
            else:   # "dv": rank r pulls rows [r*b, (r+1)*b) of every rank's contributions
                outs = [torch.zeros(b, P, dtype=torch.float64) for _ in range(world)]
                ins = [blk.acc_v[q * b:(q + 1) * b].contiguous() for q in range(world)]
                dist.all_to_all(outs, ins) if dist.get_backend() != "gloo" else _a2a_gloo(outs, ins, rank, world)
                blk.dv_parts.copy_(torch.stack(outs))

        sym = RowBlockItc(blk, exchange, b, world, rank)
        sums = torch.zeros(2, dtype=torch.float64)
        sym.forward(T, V, V_all, scale, sums)
        dT, dV = torch.empty(b, P, dtype=torch.float64), torch.empty(b, P, dtype=torch.float64)
        rsum = torch.zeros(1, dtype=torch.float64)
        sym.backward(T, V, V_all, scale, 1.0, dT_f32=dT, dV_f32=dV, r_sum=rsum)
        dist.all_reduce(sums)
        dist.all_reduce(rsum)
        loss = 0.5 * (sums[0] + sums[1]) / N
        Tq, Vq = T_full.clone().requires_grad_(True), V_full.clone().requires_grad_(True)
        ls = torch.tensor(2.6592, dtype=torch.float64, requires_grad=True)
        ref = R.clip_loss(R.itc_logits(Tq, Vq, ls))
        ref.backward()
        ok = (abs(float(loss) - float(ref)) < 1e-10
              and torch.allclose(dT, Tq.grad[rank * b:(rank + 1) * b], rtol=1e-8, atol=1e-12)
              and torch.allclose(dV, Vq.grad[rank * b:(rank + 1) * b], rtol=1e-8, atol=1e-12)
              and abs(float(rsum) - float(ls.grad)) < 1e-9)
        out_q.put((rank, bool(ok), float(loss), float(ref)))
    finally:
        dist.destroy_process_group()


def _a2a_gloo(outs, ins, rank, world):
    """all_to_all for gloo (which lacks it): rank q receives ins[q] of every rank, via one gather per destination."""
    for dst in range(world):
        gl = [torch.zeros_like(ins[dst]) for _ in range(world)] if rank == dst else None
        dist.gather(ins[dst], gl, dst=dst)
        if rank == dst:
            for q in range(world):
                outs[q].copy_(gl[q])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("worker,b,P", [(_worker, 8, 16), (_worker, 40, 64), (_worker_rowblock, 8, 16), (_worker_rowblock, 40, 64)])
def test_peer_itc_matches_oracle_world2(worker, b, P):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, b, P, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=120) for _ in procs]
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    assert all(ok for _, ok, _, _ in res), res


@pytest.mark.parametrize("push", ["beside", "before"])
def test_peer_itc_push_form_sequencing_world2(push):
    """The PUSH form of the symmetric exchange (SymmetricItc(push=...)): same losses and gradients as the oracle, the
    pull-form exchange is never called, the embeddings are pushed once, the lse vectors ride on lse_rows, every tile call
    receives the flag tuple of its phase, and `done` is signalled exactly once after the last reader."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 8, 16, q, push)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=120) for _ in procs]
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    assert all(ok for _, ok, _, _ in res), res
