"""CPU, world_size 2 over gloo: the collective sequencing of the row-sharded ITC step (tic_b200.dist.ShardedItc —
all_gather, SUM all-reduce of the column partial sums, reduce-scatter of the image-side gradient) reproduces the
single-process oracle on the concatenated batch.  The CUDA kernels need a GPU, so the per-rank block backend here is a
torch stand-in with the same piece interface as plan.ItcPlan (the GPU tests cover the kernels themselves)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restatement as R


class CpuItcBlock:
    """Stand-in for plan.ItcPlan on CPU tensors (fp64): same pieces, same meaning of every buffer."""

    def __init__(self, m, n, P, row_offset):
        self.m, self.n, self.P, self.row_offset = m, n, P, row_offset

    def norms(self, T, ldt, V, ldv, T_lo=None, V_lo=None, t_only=False):
        self.rinv_t = 1.0 / T.norm(dim=1)
        self.rinv_v = 1.0 / V.norm(dim=1)

    def fwd_tiles(self, T, ldt, V, ldv, scale, T_lo=None, V_lo=None):
        self.S = scale * (T * self.rinv_t[:, None]) @ (V * self.rinv_v[:, None]).t()
        E = torch.exp(self.S - scale)
        self.row_sum, self.col_part = E.sum(1), E.sum(0)
        self.diag = self.S[torch.arange(self.m), self.row_offset + torch.arange(self.m)]

    def reduce_col_parts(self):
        return self.col_part.clone()

    def lse_loss(self, scale, loss_sums, col_parts=None, n_col_parts=None):
        self.lse_row = scale + torch.log(self.row_sum)
        self.lse_col = scale + torch.log(col_parts)
        loss_sums[0] += (self.lse_row - self.diag).sum()
        loss_sums[1] += (self.lse_col[self.row_offset:self.row_offset + self.m] - self.diag).sum()

    def bwd_operands(self, T, ldt, V, ldv, scale, gscale, T_lo=None, V_lo=None):
        Gp = gscale * (torch.exp(self.S - self.lse_row[:, None]) + torch.exp(self.S - self.lse_col[None, :]))
        self.GA, self.GBT = Gp * self.rinv_v[None, :], (Gp * self.rinv_t[:, None]).t()

    def grad_gemms(self, T, ldt, V, ldv, T_lo=None, V_lo=None):
        self.acc_t, self.acc_v = self.GA @ V, self.GBT @ T

    @staticmethod
    def _finalize(acc, X, rinv, Xo, rinv_o, scale, diag_coef):
        dxh = scale * (acc - diag_coef * rinv_o[:, None] * Xo)
        xh = X * rinv[:, None]
        r = (xh * dxh).sum(1)
        return rinv[:, None] * (dxh - xh * r[:, None]), r

    def finalize_t(self, T, ldt, V_diag, ldv, rinv_v_diag, scale, diag_coef, dT_f32, dT_bf16, r_sum, **kw):
        g, r = self._finalize(self.acc_t, T, self.rinv_t, V_diag, rinv_v_diag, scale, diag_coef)
        dT_f32.copy_(g)
        if r_sum is not None:
            r_sum += r.sum()

    def finalize_v(self, acc_v, V, ldv, rinv_v, T_diag, ldt, rinv_t_diag, rows, scale, diag_coef, dV_f32, dV_bf16, **kw):
        g, _ = self._finalize(acc_v, V, rinv_v, T_diag, rinv_t_diag, scale, diag_coef)
        dV_f32.copy_(g)


def _worker(rank, world, port, b, P, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tic_b200.dist import ShardedItc
        N = b * world
        g = torch.Generator().manual_seed(5)
        T_all = torch.randn(N, P, generator=g, dtype=torch.float64)
        V_all = torch.randn(N, P, generator=g, dtype=torch.float64) + 0.5 * T_all
        T, V = T_all[rank * b:(rank + 1) * b].clone(), V_all[rank * b:(rank + 1) * b].clone()
        scale = float(np.exp(2.6592))
        blk = CpuItcBlock(b, N, P, rank * b)
        sh = ShardedItc(blk, b, world, rank, P)
        sums = torch.zeros(2, dtype=torch.float64)
        sh.forward(T, V, scale, sums)
        dT, dV = torch.empty(b, P, dtype=torch.float64), torch.empty(b, P, dtype=torch.float64)
        rsum = torch.zeros(1, dtype=torch.float64)
        sh.backward(T, V, scale, 1.0, dT_f32=dT, dV_f32=dV, r_sum=rsum)
        dist.all_reduce(sums)
        dist.all_reduce(rsum)
        loss = 0.5 * (sums[0] + sums[1]) / N
        # single-process oracle on the whole batch
        Tq, Vq = T_all.clone().requires_grad_(True), V_all.clone().requires_grad_(True)
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


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("b,P", [(6, 16), (32, 64)])
def test_sharded_itc_matches_oracle_world2(b, P):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, b, P, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=120) for _ in procs]
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    assert all(ok for _, ok, _, _ in res), res
