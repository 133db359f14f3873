"""Peer-memory (NVLink 5 / NVSwitch) execution of the row-sharded ITC step: no collective library on the data path.

The reference is single-device (models/mm_late.py:30); SURVEY.md §8(e) shards ITC by batch.  The NCCL sequencing lives in
`dist.ShardedItc` (all_gather -> tiles -> all_reduce of column sums -> reduce_scatter of dV).  Here the same step is
restated so that NOTHING has to be reduced across ranks:

    rank r owns text rows and image rows [r*b, (r+1)*b) of the global batch N = b * world, and publishes them (bf16
    embeddings + inverse norms) in a block of device memory that every peer maps through CUDA IPC.

    exchange #1  (one kernel: system-scope barrier + pull over NVLink)   T_all, V_all, rinv_t_all, rinv_v_all
    row block    S  [rows_r, :]  = scale * T_r V_all^T   -> row sums    = softmax statistics of my TEXT rows
    swapped      S^T[cols_r, :]  = scale * V_r T_all^T   -> row sums    = softmax statistics of my IMAGE columns
    lse + loss terms of my rows/columns (fixed-order, deterministic)
    exchange #2  lse_row_all, lse_col_all   (N floats each)
    row block    GA  = G'[rows_r, :] * rinv_v  ->  dT_r = GA  V_all          G' = g/(2N) (e^{S-lse_row} + e^{S-lse_col})
    swapped      GA' = G'^T[cols_r, :] * rinv_t -> dV_r = GA' T_all
    finalise dT_r, dV_r locally (diagonal term, normalise-backward, d logit_scale)

Both softmax directions are complete on the rank that owns them, and so are both gradients: compared with the NCCL form
this trades the all-reduce and the [N, P] fp32 reduce-scatter for recomputing the swapped tile block (tensor-core time,
which the step has to spare) and leaves two tiny exchanges per step.  Everything is enqueued on CUDA streams — the whole
multi-GPU step replays as one CUDA graph per rank.

`SymmetricItc` is the backend-agnostic sequencing (the gloo/CPU test plugs in torch stand-ins); `PeerGroup` is the
IPC plumbing + exchange kernel binding; `PeerHeadPlan` is the multi-GPU HeadPlan built on both.
"""
import ctypes
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


class SymmetricItc:
    """Sequencing of the symmetric row/column-block ITC step.

    rb / cb: the row block and the swapped block.  Each implements the `plan.ItcPlan` piece interface and carries
        rinv_t [b] (its A rows: mine), rinv_v [N] (its B rows: gathered), lse_row [b] (mine), lse_col [N] (gathered).
    exchange(phase): phase "emb" fills T_all/V_all (+ residuals) and the gathered inverse norms from every rank's
        published block; phase "lse" fills the gathered lse vectors.
    lse_rows(rb, cb, scale, loss_sums): row partials of both blocks -> rb.lse_row, cb.lse_row (published), loss terms.
    """

    def __init__(self, rb, cb, exchange, lse_rows, b_local: int, world: int, rank: int, branches=None, push=None):
        """push (optional): {"emb": callable, "seg_emb": seg tuple, "lse": callable, "seg_lse": seg tuple, "done": callable} — the
        PUSH form of the two exchanges (tic_peer_push): the producer stores into every peer's gathered buffers and releases a
        flag, the tile kernels wait per column segment themselves (seg tuples as in ItcPlan.fwd_tiles); "done" signals the
        peers, after the last read of the gathered embeddings, that the next step may overwrite them."""
        self.push = push
        self.rb, self.cb, self.exchange, self.lse_rows = rb, cb, exchange, lse_rows
        # residual (lo) K-segment of the GATHERED embeddings in the two gradient GEMMs: kept while the sum is short; from 1024
        # gathered rows on its rounding averages out (K = N terms) and the segment is a third of a long k-loop (N = 2048: 23 us)
        self.gemm_lo = b_local * world < 1024
        self.b, self.world, self.rank, self.N = b_local, world, rank, b_local * world
        # The push kernel may run BESIDE the tile kernels that poll the peers' flags only while those kernels cannot occupy
        # every SM: a polling CTA holds its SM (whole register file) until the peers' pushes land, and the peers' pushes need
        # their own SMs the same way — with both ranks' machines full of polling CTAs nobody pushes (seen as an intermittent
        # hang at b = 1024 on 2 GPUs).  Bound: CTAs of the two forward tile kernels (128 x 64 tiles) <= 132 of the 148 SMs;
        # above it the push is stream-ordered in front of them (its ~10 us are small against a step of that size).
        self.push_beside = 2 * (-(-b_local // 128)) * (-(-(b_local * world) // 64)) <= 132
        # the row block and the swapped block are independent between the exchanges: issue them on parallel branches
        self.br = branches if branches is not None else _NoBranches()

    def forward(self, T, V, T_all, V_all, scale, loss_sums, T_lo=None, V_lo=None, T_all_lo=None, V_all_lo=None,
                produce_t=None, produce_v=None):
        """produce_t / produce_v: optional callables that compute this rank's T / V into the published block (e.g. the
        projection GEMMs); they run on the two branches so that neither tower waits for the other."""
        rb, cb, br = self.rb, self.cb, self.br
        ldt, ldv = T.stride(0), V.stride(0)
        with br("cb"):
            if produce_v is not None:
                produce_v()
            cb.norm_t(V, ldv, T_lo=V_lo)                  # -> published inverse norms of my image rows
        if produce_t is not None:
            produce_t()
        rb.norm_t(T, ldt, T_lo=T_lo)                      # -> ... of my text rows
        br.join("cb")
        pu = self.push
        seg = pu["seg_emb"] if pu else None
        if pu and self.push_beside:
            with br("px"):            # my rows -> every REMOTE rank's gathered buffers + flag, on a side branch: the tile kernels
                pu["emb"]()           # below do not wait for this kernel, they poll the peers' flags segment by segment
        elif pu:
            pu["emb"]()               # large tile grids: the push is a predecessor of the polling kernels (see push_beside)
        else:
            self.exchange("emb")
        with br("cb"):
            cb.fwd_tiles(V, ldv, T_all, T_all.stride(0), scale, T_lo=V_lo, V_lo=T_all_lo, seg=seg)
        rb.fwd_tiles(T, ldt, V_all, V_all.stride(0), scale, T_lo=T_lo, V_lo=V_all_lo, seg=seg)
        br.join("cb")
        self.lse_rows(rb, cb, scale, loss_sums)
        if pu:
            with br("pl"):
                pu["lse"]()
        else:
            self.exchange("lse")

    def backward(self, T, V, T_all, V_all, scale, g, dT_f32=None, dT_bf16=None, dV_f32=None, dV_bf16=None, r_sum=None,
                 T_lo=None, V_lo=None, T_all_lo=None, V_all_lo=None, dT_lo=None, dV_lo=None, consume_t=None, consume_v=None):
        """g = dLoss/d(clip_loss) of the GLOBAL loss; produces this rank's dT, dV [b, P] and its share of d logit_scale.
        consume_t / consume_v: optional callables issued right after dT / dV exist (projection weight gradients)."""
        rb, cb, N, br = self.rb, self.cb, self.N, self.br
        ldt, ldv = T.stride(0), V.stride(0)
        pu = self.push
        seg = pu["seg_lse"] if pu else None
        ev_cb = None
        # split-K gradient GEMMs accumulate with atomics: their accumulators are zeroed on side branches beside the tile kernels
        # (in stream order the two ~2 us memsets sat between the tiles and the GEMMs of the step's critical chain)
        zc = getattr(cb, "splitk_grad", False) and getattr(br, "enabled", False)
        zr = getattr(rb, "splitk_grad", False) and getattr(br, "enabled", False)
        if zc:
            with br("zc"):
                cb.acc_t.zero_()
        if zr:
            with br("zr"):
                rb.acc_t.zero_()
        with br("cb"):
            cb.bwd_operands(V, ldv, T_all, T_all.stride(0), scale, g / (2.0 * N), T_lo=V_lo, V_lo=T_all_lo, seg=seg)
            if zc:
                br.join("zc")
                cb.grad_gemm_t(T_all, T_all.stride(0), V_lo=T_all_lo if self.gemm_lo else None, prezeroed=True)
            else:
                cb.grad_gemm_t(T_all, T_all.stride(0), V_lo=T_all_lo if self.gemm_lo else None)
            if pu and T.is_cuda:
                ev_cb = torch.cuda.Event()
                ev_cb.record(torch.cuda.current_stream())
            cb.finalize_t(V, ldv, T, ldt, rb.rinv_t, scale, g / N, dV_f32, dV_bf16, None, dT_lo=dV_lo, T_lo=V_lo, V_diag_lo=T_lo)
            if consume_v is not None:
                consume_v()
        rb.bwd_operands(T, ldt, V_all, V_all.stride(0), scale, g / (2.0 * N), T_lo=T_lo, V_lo=V_all_lo, seg=seg)
        if zr:
            br.join("zr")
            rb.grad_gemm_t(V_all, V_all.stride(0), V_lo=V_all_lo if self.gemm_lo else None, prezeroed=True)
        else:
            rb.grad_gemm_t(V_all, V_all.stride(0), V_lo=V_all_lo if self.gemm_lo else None)
        if pu:      # both gradient GEMMs were the last readers of the gathered embeddings: tell the peers (side branch)
            with br("dn"):
                if ev_cb is not None:
                    torch.cuda.current_stream().wait_event(ev_cb)
                pu["done"]()
        rb.finalize_t(T, ldt, V, ldv, cb.rinv_t, scale, g / N, dT_f32, dT_bf16, r_sum, dT_lo=dT_lo, T_lo=T_lo, V_diag_lo=V_lo)
        if consume_t is not None:
            consume_t()
        br.join("cb")
        if pu:
            br.join("dn")
            br.join("px")
            br.join("pl")


class RowBlockItc:
    """Sequencing of the row-block ITC step with a peer-memory reduction of the image-side gradient — the form used from
    4096 global negatives on, where recomputing the swapped tile block (SymmetricItc) would cost more tensor time than
    moving the gradient contributions:

        exchange "emb"   barrier + rinv_v_all; then V_all is pulled peer by peer on a side branch while the row-block tile
                         kernel already runs: its TMA producer waits per segment (local segment first)  — the all-gather
                         is consumed as it lands                       (text rows never leave their rank)
        row block        S[rows_r, :] -> row sums (complete), column partial sums [N] (published)
        exchange "col"   every rank's column sums [R, N] -> summed in rank order -> lse_col (identical on all ranks)
        row block        GA = G'[rows_r, :] * rinv_v ;  dT_r = GA V_all (local)
                         dV contributions  GA^T That_r  [N, P] fp32 (published)
        exchange "dv"    rows [r*b, (r+1)*b) of every rank's contributions [R, b, P] -> summed in rank order -> dV_r

    blk implements the plan.ItcPlan piece interface (GA-shared mode); publish_v_norm / publish_col_sums / lse_loss_gathered
    / reduce_dv are the pieces that touch the published block."""

    def __init__(self, blk, exchange, b_local: int, world: int, rank: int, branches=None, pull=None, seg=None):
        """pull(name): the barrier-free pull half of an exchange, issued on a side branch; seg: the (ready, epoch, seg_cols,
        my_seg) description handed to fwd_tiles so that the tiles consume V_all segment by segment while it lands.  Without
        `pull` the "emb" exchange gathers V_all itself (gloo/CPU stand-in)."""
        self.blk, self.exchange, self.pull, self.seg = blk, exchange, pull, seg
        import os as _os
        self.overlap = _os.environ.get("TIC_PEER_OVERLAP", "1") != "0"
        self.b, self.world, self.rank, self.N = b_local, world, rank, b_local * world
        self.br = branches if branches is not None else _NoBranches()

    def forward(self, T, V, V_all, scale, loss_sums, produce_t=None, produce_v=None):
        blk, br = self.blk, self.br
        ldt = T.stride(0)
        with br("cb"):
            if produce_v is not None:
                produce_v()
            blk.publish_v_norm(V)                               # inverse norms of my image rows -> published
        if produce_t is not None:
            produce_t()
        blk.norm_t(T, ldt)                                      # local (+ normalised bf16 copy for the dV product)
        br.join("cb")
        self.exchange("emb")                                    # barrier (+ inverse norms; + V_all when there is no pull)
        if self.pull is not None and self.overlap:
            with br("pull"):
                self.pull("vall")                               # V_all lands peer by peer beside the tile kernel
            blk.fwd_tiles(T, ldt, V_all, V_all.stride(0), scale, seg=self.seg)
            br.join("pull")
        elif self.pull is not None:                             # TIC_PEER_OVERLAP=0: pull, then tiles (A/B switch)
            self.pull("vall")
            blk.fwd_tiles(T, ldt, V_all, V_all.stride(0), scale)
        else:
            blk.fwd_tiles(T, ldt, V_all, V_all.stride(0), scale)
        blk.publish_col_sums()
        self.exchange("col")
        blk.lse_loss_gathered(scale, loss_sums)

    def backward(self, T, V, V_all, scale, g, dT_f32=None, dT_bf16=None, dV_f32=None, dV_bf16=None, r_sum=None, dT_lo=None,
                 dV_lo=None, consume_t=None, consume_v=None):
        blk, N, b, br = self.blk, self.N, self.b, self.br
        ldt, ldv = T.stride(0), V.stride(0)
        blk.bwd_operands(T, ldt, V_all, V_all.stride(0), scale, g / (2.0 * N))
        with br("cb"):   # image side first: its exchange overlaps the text-side GEMM on the main stream
            blk.grad_gemm_v(T, ldt)                             # -> published contributions [N, P]
            self.exchange("dv")
            acc_v = blk.reduce_dv()                             # [b, P], fixed rank order
            blk.finalize_v(acc_v, V, ldv, blk.rinv_v_mine(), T, ldt, blk.rinv_t, b, scale, g / N, dV_f32, dV_bf16, dV_lo=dV_lo)
            if consume_v is not None:
                consume_v()
        blk.grad_gemm_t(V_all, V_all.stride(0))
        blk.finalize_t(T, ldt, V, ldv, blk.rinv_v_mine(), scale, g / N, dT_f32, dT_bf16, r_sum, dT_lo=dT_lo)
        if consume_t is not None:
            consume_t()
        br.join("cb")


class _NoBranches:
    class _Ctx:
        def __enter__(self):
            return None

        def __exit__(self, *a):
            return False

    def __call__(self, k):
        return self._Ctx()

    def join(self, k):
        pass


class _RawCuda:
    """Exposes a raw device allocation through __cuda_array_interface__ so torch can view it without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def _up(x, a=256):
    return (x + a - 1) // a * a


class PeerGroup:
    """One symmetric block of `nbytes` per rank, mapped by every peer.  Handles are exchanged once through the process
    group (setup only); `exchange()` is the per-step kernel and touches no collective library."""

    def __init__(self, nbytes: int, world: int, rank: int, device, group=None):
        from . import capi
        lib = capi.load()
        assert 1 <= world <= 8, "tic_peer_exchange supports up to 8 ranks (one NVSwitch domain)"
        self.capi, self.world, self.rank, self.dev = capi, world, rank, torch.device(device)
        self.nbytes = _up(nbytes)
        self.flag_off = self.nbytes
        total = self.nbytes + 256
        p = ctypes.c_void_p()
        capi.call("tic_peer_alloc", total, ctypes.byref(p))
        self.local = int(p.value)
        hb = lib.tic_peer_handle_bytes()
        buf = (ctypes.c_ubyte * hb)()
        capi.call("tic_peer_export", self.local, buf)
        mine = torch.tensor(list(buf), dtype=torch.uint8, device=self.dev)
        allh = torch.empty(world * hb, dtype=torch.uint8, device=self.dev)
        dist.all_gather_into_tensor(allh, mine, group=group)
        allh = allh.cpu().numpy().reshape(world, hb)
        self.bases: List[int] = []
        for r in range(world):
            if r == rank:
                self.bases.append(self.local)
                continue
            hbuf = (ctypes.c_ubyte * hb)(*[int(x) for x in allh[r]])
            q = ctypes.c_void_p()
            capi.call("tic_peer_open", hbuf, ctypes.byref(q))
            self.bases.append(int(q.value))
        self._bases_c = (ctypes.c_void_p * world)(*self.bases)
        self.ctr = torch.zeros(2, dtype=torch.int32, device=self.dev)
        self._holder = _RawCuda(self.local, total)
        self.block = torch.as_tensor(self._holder, device=self.dev)
        assert self.block.data_ptr() == self.local, "torch copied the peer block instead of viewing it"
        self._phases = {}
        self.define_phase("align", [])     # barrier only
        torch.cuda.synchronize(self.dev)
        dist.barrier(group=group)     # every block is allocated, zeroed and mapped before the first exchange

    def view(self, off: int, shape: Tuple[int, ...], dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        nb = n * torch.empty(0, dtype=dtype).element_size()
        assert off % 256 == 0 and off + nb <= self.nbytes
        return self.block[off:off + nb].view(dtype).view(*shape)

    def define_phase(self, name: str, segs):
        """segs: [(src_off_bytes, nbytes, dst_tensor, dst_stride_bytes)] — dst receives rank p's range at p*dst_stride."""
        n = len(segs)
        self._phases[name] = (n, (ctypes.c_int64 * n)(*[s[0] for s in segs]), (ctypes.c_int64 * n)(*[s[1] for s in segs]),
                              (ctypes.c_void_p * n)(*[s[2].data_ptr() for s in segs]),
                              (ctypes.c_int64 * n)(*[s[3] for s in segs]), [s[2] for s in segs])

    def exchange(self, name: str):
        n, so, nb, dst, ds, _keep = self._phases[name]
        self.capi.call("tic_peer_exchange", self._bases_c, self.world, self.rank, self.flag_off, self.ctr.data_ptr(), n, so, nb,
                       dst, ds, torch.cuda.current_stream().cuda_stream)

    # flag words of the push form: uint32[8] slots inside the 256-byte flag area of every block (slot 0 = the pull barrier)
    FLAG_SLOT_BYTES = 32

    def flags_view(self, slot: int) -> torch.Tensor:
        """the LOCAL uint32[world] flag words of a slot (written by the peers), as an int32 tensor"""
        off = self.flag_off + slot * self.FLAG_SLOT_BYTES
        return self.block[off:off + 4 * self.world].view(torch.int32)

    def step_counter(self) -> torch.Tensor:
        """{completed steps, epoch of the step in flight} (tic_peer_signal advances it at the tail of a step)"""
        if getattr(self, "_step", None) is None:
            self._step = torch.tensor([0, 1], dtype=torch.int32, device=self.dev)
        return self._step

    def define_push(self, name: str, segs, slot: int, wait_slot: int = -1):
        """segs: [(src_tensor = this rank's slot of its own gathered buffer, nbytes, dst_off_bytes of rank 0's slot in every
        block, dst_stride_bytes)]"""
        n = len(segs)
        ticket = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._pushes = getattr(self, "_pushes", {})
        self._pushes[name] = (n, (ctypes.c_void_p * n)(*[s[0].data_ptr() for s in segs]), (ctypes.c_int64 * n)(*[s[1] for s in segs]),
                              (ctypes.c_int64 * n)(*[s[2] for s in segs]), (ctypes.c_int64 * n)(*[s[3] for s in segs]),
                              self.flag_off + slot * self.FLAG_SLOT_BYTES,
                              (self.flag_off + wait_slot * self.FLAG_SLOT_BYTES) if wait_slot >= 0 else -1, ticket, [s[0] for s in segs])

    def push(self, name: str):
        n, src, nb, doff, dstr, foff, woff, ticket, _keep = self._pushes[name]
        self.capi.call("tic_peer_push", self._bases_c, self.world, self.rank, foff, woff, self.step_counter().data_ptr(),
                       ticket.data_ptr(), n, src, nb, doff, dstr, torch.cuda.current_stream().cuda_stream)

    def define_signal(self, name: str, slot: int):
        self._signals = getattr(self, "_signals", {})
        self._signals[name] = self.flag_off + slot * self.FLAG_SLOT_BYTES

    def signal(self, name: str):
        self.capi.call("tic_peer_signal", self._bases_c, self.world, self.rank, self._signals[name], self.step_counter().data_ptr(),
                       torch.cuda.current_stream().cuda_stream)

    def define_pull(self, name: str, segs, max_blocks: int = 148):
        n = len(segs)
        ready = torch.zeros(self.world, dtype=torch.int32, device=self.dev)
        tickets = torch.zeros(self.world, dtype=torch.int32, device=self.dev)
        self._pulls = getattr(self, "_pulls", {})
        self._pulls[name] = (n, (ctypes.c_int64 * n)(*[s[0] for s in segs]), (ctypes.c_int64 * n)(*[s[1] for s in segs]),
                             (ctypes.c_void_p * n)(*[s[2].data_ptr() for s in segs]),
                             (ctypes.c_int64 * n)(*[s[3] for s in segs]), ready, tickets, max_blocks, [s[2] for s in segs])
        return ready

    def pull(self, name: str):
        n, so, nb, dst, ds, ready, tickets, mb, _keep = self._pulls[name]
        self.capi.call("tic_peer_pull", self._bases_c, self.world, self.rank, self.ctr.data_ptr(), ready.data_ptr(),
                       tickets.data_ptr(), n, so, nb, dst, ds, mb, torch.cuda.current_stream().cuda_stream)

    def close(self):
        for r, b in enumerate(self.bases):
            if r != self.rank and b:
                self.capi.call("tic_peer_close", b)
        self.bases = []


def _make_peer_head_plan():
    from . import plan as P
    from .capi import call, ptr

    BF16, F32 = torch.bfloat16, torch.float32

    class PeerHeadPlan(P.HeadPlan):
        """Multi-GPU HeadPlan over peer memory: ITC is row-sharded with the symmetric formulation above; fusion heads, ITM
        sampling/gather and the small heads are per-sample (pure data parallel).  Every rank produces the gradients of the
        GLOBAL loss restricted to its samples (weight gradients / d logit_scale are summed across ranks by the caller)."""

        def __init__(self, B_local, *, world, rank, group=None, d: Optional[int] = None, itc_mode: Optional[str] = None, **kw):
            kw.setdefault("use_itc", True)
            use_itc = kw["use_itc"]
            kw["use_itc"] = False                       # the parent must not allocate a square single-GPU ItcPlan
            super().__init__(B_local, **kw)
            self.use_itc = use_itc
            self.world, self.rank = world, rank
            b, dev = B_local, self.dev
            if self.P is None and d is not None:
                self.Pe = d
                self.out["d_t_emb"] = torch.empty(b, d, device=dev)
                self.out["d_v_emb"] = torch.empty(b, d, device=dev)
            N = b * world
            self.n_global = N
            if use_itc:
                self.beta_itc = kw.get("beta_itc", 0.1) if self.fusion is not None else 1.0
                self.g_itc = self.beta_itc if self.fusion is not None else 1.0
                self.w_cls = (1.0 - (self.beta_itc + self.beta_itm)) if self.fusion is not None else 0.0
            self.w_cls /= world                          # local means -> contributions to global means
            self.beta_itm_local = self.beta_itm / world
            if not use_itc:
                return
            assert b % 8 == 0, "per-rank batch must be a multiple of 8 (16-byte exchange segments)"
            # symmetric blocks while the step is latency-bound; row block + peer reduction once tensor time dominates
            import os as _os
            itc_mode = itc_mode or _os.environ.get("TIC_PEER_ITC_MODE") or None     # env: A/B measurement switch
            self.itc_mode = itc_mode if itc_mode is not None else ("symmetric" if N < 4096 else "rowblock")
            assert self.itc_mode in ("symmetric", "rowblock")
            if self.itc_mode == "symmetric":
                self._init_symmetric(b, N, group)
            else:
                self._init_rowblock(b, N, group)

        def _init_symmetric(self, b, N, group):
            world, rank, dev, Pe = self.world, self.rank, self.dev, self.Pe
            precise = N < 4096
            self.has_lo = precise and self.P is not None  # residuals exist only for embeddings projected on the device
            ybytes = 2 * b * Pe * 2
            off_hi, off_lo = 0, _up(ybytes)
            off_rinv = off_lo + (_up(ybytes) if self.has_lo else 0)
            off_lse = off_rinv + _up(2 * b * 4)
            end_pub = off_lse + _up(2 * b * 4)
            import os as _os
            # PUSH form of the two exchanges (default): the gathered copies live INSIDE the peer-mapped block so that the
            # producers can store into them; TIC_PEER_PUSH=0 keeps the round-1 pull form (A/B switch)
            self.push_mode = _os.environ.get("TIC_PEER_PUSH", "1") != "0"
            self._lse_push = None
            g_emb = _up(N * Pe * 2)
            off_Tall, off_Vall = end_pub, end_pub + g_emb
            off_Tall_lo, off_Vall_lo = off_Vall + g_emb, off_Vall + 2 * g_emb
            off_rt_all = off_Vall + (3 * g_emb if self.has_lo else g_emb)
            off_rv_all, off_lr_all, off_lc_all = off_rt_all + _up(N * 4), off_rt_all + 2 * _up(N * 4), off_rt_all + 3 * _up(N * 4)
            total = (off_rt_all + 4 * _up(N * 4)) if self.push_mode else end_pub
            self.pg = PeerGroup(total, world, rank, dev, group)
            pg = self.pg
            e = lambda *s, dt=F32: torch.empty(*s, dtype=dt, device=dev)  # noqa: E731
            # published (symmetric) views: [text rows ; image rows]
            self.Y = pg.view(off_hi, (2 * b, Pe), BF16)
            self.Y_lo = pg.view(off_lo, (2 * b, Pe), BF16) if self.has_lo else (e(2 * b, Pe, dt=BF16) if self.P is not None else None)
            self.rinv_mine = pg.view(off_rinv, (2 * b,), F32)
            self.lse_mine = pg.view(off_lse, (2 * b,), F32)
            self._Yt = self._Yv = self._Yt_lo = self._Yv_lo = None
            if self.push_mode:   # gathered copies inside the peer-mapped block (the peers write them)
                self.T_all, self.V_all = pg.view(off_Tall, (N, Pe), BF16), pg.view(off_Vall, (N, Pe), BF16)
                self.T_all_lo = pg.view(off_Tall_lo, (N, Pe), BF16) if self.has_lo else None
                self.V_all_lo = pg.view(off_Vall_lo, (N, Pe), BF16) if self.has_lo else None
                self.rinv_t_all, self.rinv_v_all = pg.view(off_rt_all, (N,), F32), pg.view(off_rv_all, (N,), F32)
                self.lse_row_all, self.lse_col_all = pg.view(off_lr_all, (N,), F32), pg.view(off_lc_all, (N,), F32)
                # this rank's rows are produced IN PLACE in its slot of its own gathered buffers (no local copy, and the local
                # column segment of the tile kernels is ready by stream order)
                sl = slice(rank * b, (rank + 1) * b)
                self._Yt, self._Yv = self.T_all[sl], self.V_all[sl]
                if self.has_lo:
                    self._Yt_lo, self._Yv_lo = self.T_all_lo[sl], self.V_all_lo[sl]
            else:                # gathered copies in local HBM (pulled)
                self.T_all, self.V_all = e(N, Pe, dt=BF16), e(N, Pe, dt=BF16)
                self.T_all_lo = e(N, Pe, dt=BF16) if self.has_lo else None
                self.V_all_lo = e(N, Pe, dt=BF16) if self.has_lo else None
                self.rinv_t_all, self.rinv_v_all, self.lse_row_all, self.lse_col_all = e(N), e(N), e(N), e(N)
            half = b * Pe * 2
            segs = [(off_hi, half, self.T_all, half), (off_hi + half, half, self.V_all, half)]
            if self.has_lo:
                segs += [(off_lo, half, self.T_all_lo, half), (off_lo + half, half, self.V_all_lo, half)]
            segs += [(off_rinv, b * 4, self.rinv_t_all, b * 4), (off_rinv + b * 4, b * 4, self.rinv_v_all, b * 4)]
            pg.define_phase("emb", segs)
            pg.define_phase("lse", [(off_lse, b * 4, self.lse_row_all, b * 4), (off_lse + b * 4, b * 4, self.lse_col_all, b * 4)])
            mk = lambda: P.ItcPlan(b, N, Pe, dev, row_offset=rank * b, need_dv=False, precise=precise, col_sums=False)  # noqa: E731
            self.rb, self.cb = mk(), mk()
            if self.push_mode:
                sl = slice(rank * b, (rank + 1) * b)
                self.rb.rinv_t, self.rb.rinv_v = self.rinv_t_all[sl], self.rinv_v_all
                self.cb.rinv_t, self.cb.rinv_v = self.rinv_v_all[sl], self.rinv_t_all
                self.rb.lse_row, self.rb.lse_col = self.lse_row_all[sl], self.lse_col_all
                self.cb.lse_row, self.cb.lse_col = self.lse_col_all[sl], self.lse_row_all
            else:
                self.rb.rinv_t, self.rb.rinv_v = self.rinv_mine[:b], self.rinv_v_all
                self.cb.rinv_t, self.cb.rinv_v = self.rinv_mine[b:], self.rinv_t_all
                self.rb.lse_row, self.rb.lse_col = self.lse_mine[:b], self.lse_col_all
                self.cb.lse_row, self.cb.lse_col = self.lse_mine[b:], self.lse_row_all
            self.itc = self.rb
            self.rb.scale_dev = self.cb.scale_dev = self.scale_t      # exp(logit_scale) by device pointer (live weights)
            self.lse_ws = torch.zeros(int(capi_load().tic_itc_lse_rows_workspace_bytes(b)) // 4, dtype=F32, device=dev)
            push = None
            if self.push_mode:
                sl = slice(rank * b, (rank + 1) * b)
                psegs = [(self._Yt, half, off_Tall, half), (self._Yv, half, off_Vall, half)]
                if self.has_lo:
                    psegs += [(self._Yt_lo, half, off_Tall_lo, half), (self._Yv_lo, half, off_Vall_lo, half)]
                psegs += [(self.rinv_t_all[sl], b * 4, off_rt_all, b * 4), (self.rinv_v_all[sl], b * 4, off_rv_all, b * 4)]
                pg.define_push("emb", psegs, slot=1, wait_slot=3)
                pg.define_push("lse", [(self.lse_row_all[sl], b * 4, off_lr_all, b * 4), (self.lse_col_all[sl], b * 4, off_lc_all, b * 4)],
                               slot=2)
                pg.define_signal("done", slot=3)
                step = pg.step_counter()
                # seg tuples: (ready words, epoch word = step[1], -columns per segment [negative: flags written by the peers], my segment)
                # the lse exchange rides on tic_itc_lse_rows (no kernel of its own): "lse" below is a no-op
                self._lse_push = (off_lr_all, off_lc_all, pg.flag_off + 2 * pg.FLAG_SLOT_BYTES)
                push = {"emb": lambda: pg.push("emb"), "lse": lambda: None, "done": lambda: pg.signal("done"),
                        "seg_emb": (pg.flags_view(1), step[1:2], -b, rank), "seg_lse": (pg.flags_view(2), step[1:2], -b, rank)}
            self.sym = SymmetricItc(self.rb, self.cb, pg.exchange, self._lse_rows, b, world, rank, branches=self.br, push=push)

        def _init_rowblock(self, b, N, group):
            world, rank, dev, Pe = self.world, self.rank, self.dev, self.Pe
            self.has_lo = False
            off_y = 0
            off_rinv = off_y + _up(2 * b * Pe * 2)
            off_col = off_rinv + _up(b * 4)
            off_acc = off_col + _up(N * 4)
            self.pg = PeerGroup(off_acc + _up(N * Pe * 4), world, rank, dev, group)
            pg = self.pg
            e = lambda *s, dt=F32: torch.empty(*s, dtype=dt, device=dev)  # noqa: E731
            self.Y = pg.view(off_y, (2 * b, Pe), BF16)                    # [text rows ; image rows] (image half is pulled by peers)
            self.Y_lo = e(2 * b, Pe, dt=BF16) if self.P is not None else None
            self.V_all, self.rinv_v_all = e(N, Pe, dt=BF16), e(N)
            col_all, dv_parts, acc_v_mine = e(world, N), e(world, b, Pe), e(b, Pe)
            half = b * Pe * 2
            pg.define_phase("emb", [(off_rinv, b * 4, self.rinv_v_all, b * 4)])
            v_ready = pg.define_pull("vall", [(off_y + half, half, self.V_all, half)])
            pg.define_phase("col", [(off_col, N * 4, col_all, N * 4)])
            pg.define_phase("dv", [(off_acc + rank * b * Pe * 4, b * Pe * 4, dv_parts, b * Pe * 4)])
            blk = P.ItcPlan(b, N, Pe, dev, row_offset=rank * b, need_dv=True, precise=False)
            blk.rinv_v = self.rinv_v_all
            blk.acc_v = pg.view(off_acc, (N, Pe), F32)                  # published gradient contributions
            blk.col_sum = pg.view(off_col, (N,), F32)                   # published column sums
            rinv_pub = pg.view(off_rinv, (b,), F32)
            plan = self

            def publish_v_norm(V):
                call("tic_row_rnorm_bf16", ptr(V), None, V.stride(0), b, Pe, ptr(rinv_pub), None, 0, P._stream())

            def publish_col_sums():
                call("tic_reduce_parts", ptr(blk.col_part), blk.ncp, N, ptr(blk.col_sum), P._stream())

            def lse_loss_gathered(scale, loss_sums):
                blk.lse_loss(scale, loss_sums, col_parts=col_all, n_col_parts=world)

            def reduce_dv():
                call("tic_reduce_parts", ptr(dv_parts), world, b * Pe, ptr(acc_v_mine), P._stream())
                return acc_v_mine

            blk.publish_v_norm, blk.publish_col_sums, blk.lse_loss_gathered, blk.reduce_dv = \
                publish_v_norm, publish_col_sums, lse_loss_gathered, reduce_dv
            blk.rinv_v_mine = lambda: rinv_pub
            self.itc = self.rb = blk
            blk.scale_dev = self.scale_t
            self._keep = (col_all, dv_parts, acc_v_mine, rinv_pub, plan)
            self.sym = RowBlockItc(blk, pg.exchange, b, world, rank, branches=self.br, pull=pg.pull,
                                   seg=(v_ready, pg.ctr, b, rank))

        def _lse_rows(self, rb, cb, scale, loss_sums):
            if getattr(self, "_lse_push", None) is not None:     # push form: the lse exchange rides on this launch
                off_a, off_b, flag_off = self._lse_push
                pg = self.pg
                call("tic_itc_lse_rows_push", ptr(rb.row_part), ptr(cb.row_part), rb.nrp, self.B, ptr(rb.diag), float(scale),
                     ptr(rb.lse_row), ptr(cb.lse_row), ptr(loss_sums), ptr(self.lse_ws), ptr(self.scale_t), pg._bases_c, pg.world, pg.rank,
                     off_a, off_b, flag_off, pg.step_counter().data_ptr(), P._stream())
                return
            call("tic_itc_lse_rows", ptr(rb.row_part), ptr(cb.row_part), rb.nrp, self.B, ptr(rb.diag), float(scale),
                 ptr(rb.lse_row), ptr(cb.lse_row), ptr(loss_sums), ptr(self.lse_ws), ptr(self.scale_t), P._stream())

        def _heads(self, inp, dH_f32=None, forward_only=False, dz_ext=None):
            saved = self.beta_itm
            self.beta_itm = self.beta_itm_local
            try:
                super()._heads(inp, dH_f32=dH_f32, forward_only=forward_only, dz_ext=dz_ext)
            finally:
                self.beta_itm = saved

        def _rows(self):
            """(Yt, Yv, Yt_lo, Yv_lo): where this rank's projected / handed-in embeddings live"""
            b = self.B
            if getattr(self, "_Yt", None) is not None:      # push form: in place in the gathered buffers
                lo_t = self._Yt_lo if self._Yt_lo is not None else (self.Y_lo[:b] if self.Y_lo is not None else None)
                lo_v = self._Yv_lo if self._Yv_lo is not None else (self.Y_lo[b:] if self.Y_lo is not None else None)
                return self._Yt, self._Yv, lo_t, lo_v
            return self.Y[:b], self.Y[b:], (self.Y_lo[:b] if self.Y_lo is not None else None), (self.Y_lo[b:] if self.Y_lo is not None else None)

        def _ops(self):
            b = self.B
            lo = self.has_lo
            if self.itc_mode == "rowblock":
                return dict(T=self.Y[:b], V=self.Y[b:], V_all=self.V_all)
            Yt, Yv, Ytl, Yvl = self._rows()
            return dict(T=Yt, V=Yv, T_all=self.T_all, V_all=self.V_all, T_lo=Ytl if lo else None, V_lo=Yvl if lo else None,
                        T_all_lo=self.T_all_lo, V_all_lo=self.V_all_lo)

        def _itc_fwd(self, inp, with_loss=True):
            B, E, w = self.B, self.E, self.w
            if self.itc_mode == "rowblock":
                Yt, Yv, Ytl, Yvl = self.Y[:B], self.Y[B:], (self.Y_lo[:B] if self.Y_lo is not None else None), \
                    (self.Y_lo[B:] if self.Y_lo is not None else None)
            else:
                Yt, Yv, Ytl, Yvl = self._rows()
            self.br.enabled = self.parallel_streams
            self._refresh("itc")       # live weights without a root refresh launch (more than 8 matrices): W_t / W_v + exp(logit_scale)
            if self.P is not None:
                tp_, vp_ = inp["t_pool"], inp["v_pool"]
                pt = lambda: P.gemm(tp_, tp_.stride(0), 0, w["W_t"], E, 0, Yt, self.P, 1, B, self.P, E, D_lo=Ytl)  # noqa: E731  HF :265
                pv = lambda: P.gemm(vp_, vp_.stride(0), 0, w["W_v"], E, 0, Yv, self.P, 1, B, self.P, E, D_lo=Yvl)  # noqa: E731  HF :262
            else:   # embeddings handed in by the caller: publish them
                pt = lambda: Yt.copy_(inp["t_pool"])  # noqa: E731
                pv = lambda: Yv.copy_(inp["v_pool"])  # noqa: E731
            self.sym.forward(scale=self.scale, loss_sums=self.z["itc_sums"], produce_t=pt, produce_v=pv, **self._ops())

        def _itc_bwd(self, inp, dS=None):
            assert dS is None, "the sharded path implements the fused loss only"
            self._join_zero()
            B, E, w, z, o, br = self.B, self.E, self.w, self.z, self.out, self.br
            br.enabled = self.parallel_streams
            if self.P is not None:
                dYt, dYv, dYt_lo, dYv_lo = self.dY[:B], self.dY[B:], self.dY_lo[:B], self.dY_lo[B:]
                tp_, vp_ = inp["t_pool"], inp["v_pool"]

                def consume_t():
                    with br("w"):
                        P.gemm(dYt, self.P, 1, tp_, tp_.stride(0), 1, o["dW_t"], E, 0, self.P, E, B, A_lo=dYt_lo, accumulate=self._atomic["dW_t"])
                    P.gemm(dYt, self.P, 0, w["W_t"], E, 1, o["d_t_pool"], E, 0, B, E, self.P, A_lo=dYt_lo)
                    br.join("w")

                def consume_v():
                    P.gemm(dYv, self.P, 1, vp_, vp_.stride(0), 1, o["dW_v"], E, 0, self.P, E, B, A_lo=dYv_lo, accumulate=self._atomic["dW_v"])

                self.sym.backward(scale=self.scale, g=self.g_itc, dT_bf16=dYt, dV_bf16=dYv, r_sum=z["r_sum"], dT_lo=dYt_lo,
                                  dV_lo=dYv_lo, consume_t=consume_t, consume_v=consume_v, **self._ops())
            else:
                self.sym.backward(scale=self.scale, g=self.g_itc, dT_f32=o["d_t_emb"], dV_f32=o["d_v_emb"], r_sum=z["r_sum"],
                                  **self._ops())

        def global_loss(self):
            """[mix, cls, itc, itm] of the GLOBAL batch (reporting only; this is the one place a collective is used)."""
            t = self.out["loss"].clone()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            cls, itc, itm = t[1] / self.world, t[2], t[3] / self.world
            if self.fusion is None:
                return torch.stack([itc, cls, itc, itm])
            mix = (1.0 - (self.beta_itc + self.beta_itm)) * cls + self.beta_itc * itc + self.beta_itm * itm
            return torch.stack([mix, cls, itc, itm])

    def capi_load():
        from . import capi
        return capi.load()

    return PeerHeadPlan


def __getattr__(name):
    if name == "PeerHeadPlan":
        return _make_peer_head_plan()
    raise AttributeError(name)
