"""Multi-GPU check, run under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py

    ... tests/dist_gpu_check.py [--mode peer|nccl] [--graph] [--live]

--mode peer (default): tic_b200.peer.PeerHeadPlan — symmetric row/column blocks over CUDA-IPC peer memory, no collective on
the data path; --mode nccl: tic_b200.dist.DistHeadPlan — all_gather / all_reduce / reduce_scatter.  --graph replays the step
as a CUDA graph (peer mode only).
Every rank runs the plan on its shard; rank 0 additionally runs the single-GPU HeadPlan on the concatenated batch.
Checked: global loss, this rank's input gradients, and the SUM-all-reduced weight gradients against the single-GPU
gradients (1e-3 scale-relative).  Exit code 0 = pass."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import tic_b200.plan as P
    mode = "peer"
    if "--mode" in sys.argv:
        mode = sys.argv[sys.argv.index("--mode") + 1]
    use_graph = "--graph" in sys.argv
    live = "--live" in sys.argv       # peer mode: fp32 master weights refreshed by the root launch of every step (as bench.py)
    if mode == "peer":
        from tic_b200.peer import PeerHeadPlan as DistHeadPlan
    else:
        from tic_b200.dist import DistHeadPlan
    from oracle import restatement as R   # weight init + sampling rule only

    ok = True
    # concat head at a small batch (symmetric form, push exchange), the c2 shape (256 per rank), a long gathered dimension in the
    # symmetric form (1024 per rank: 2048 global, the 8-GPU c2 width), the row-block form (2048 per rank)
    for fusion, b, use_itm in (("concat", 96, True), ("concat", 256, True), ("concat", 1024, False), (None, 2048, False)):
        C, E = 4, 768
        N = b * world
        g = torch.Generator().manual_seed(11)
        full = {"t_pool": torch.tanh(torch.randn(N, E, generator=g)), "v_pool": torch.tanh(torch.randn(N, E, generator=g)),
                "x_t": torch.randn(N, 1, E, generator=g), "x_v": torch.randn(N, 1, E, generator=g),
                "y_soft": torch.eye(C)[torch.randint(0, C, (N,), generator=g)]}
        if fusion is None:
            full = {"t_pool": torch.randn(N, E, generator=g), "v_pool": torch.randn(N, E, generator=g)}
            full["v_pool"] = full["v_pool"] + 0.3 * full["t_pool"]
        # ITM decisions: in-shard negatives (the sampler is per-rank data parallel)
        rs = np.random.RandomState(3)
        lbl_all, src_all = [], []
        for r in range(world):
            l, s = R.itm_sample_uniform(rs.uniform(size=b).astype(np.float32), rs.uniform(size=b).astype(np.float32))
            lbl_all.append(l); src_all.append(s + r * b)
        bfk = ("t_pool", "v_pool", "x_t", "x_v")
        to_dev = lambda d: {k: (v.to(torch.bfloat16) if k in bfk else v).to(dev).contiguous() for k, v in d.items()}  # noqa: E731
        shard = to_dev({k: v[rank * b:(rank + 1) * b] for k, v in full.items()})
        if use_itm:
            shard["lbl_tim"] = torch.from_numpy(lbl_all[rank]).to(dev)
            shard["src_idx"] = torch.from_numpy((src_all[rank] - rank * b).astype(np.int32)).to(dev)
        w32 = R.init_params(C, seed=5)
        kw = dict(E=E, P=(512 if fusion is not None else None), C=C, fusion=fusion, use_itc=True, use_itm=use_itm, Lv=1)
        dplan = DistHeadPlan(b, world=world, rank=rank, d=(E if fusion is None else None), device=dev, **kw)
        if live and mode == "peer":
            dplan.bind_params({k: torch.as_tensor(v).to(dev) for k, v in w32.items()}, live=True)
        else:
            dplan.set_weights(w32)
        out = dplan.step(shard)
        torch.cuda.synchronize()
        if use_graph:      # the whole multi-GPU step (exchanges included) as one CUDA graph, replayed twice
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                dplan.step(shard)
            torch.cuda.current_stream().wait_stream(side)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                dplan.step(shard)
            for _ in range(2):
                gr.replay()
            torch.cuda.synchronize()
            out = dplan.out
        gl = dplan.global_loss()
        wkeys = [k for k in out if k.startswith("dW_") or k.startswith("db_")] + ["d_logit_scale"]
        wsum = {}
        for k in wkeys:
            t = out[k].clone().float()
            dist.all_reduce(t)
            wsum[k] = t
        if rank == 0:
            splan = P.HeadPlan(N, device=dev, **kw)
            if fusion is None:
                splan.itc = P.ItcPlan(N, N, E, dev)
                splan.Pe = E
                splan.out["d_t_emb"], splan.out["d_v_emb"] = torch.empty(N, E, device=dev), torch.empty(N, E, device=dev)
            splan.set_weights(w32)
            fin = to_dev(full)
            if use_itm:
                fin["lbl_tim"] = torch.from_numpy(np.concatenate(lbl_all)).to(dev)
                fin["src_idx"] = torch.from_numpy(np.concatenate(src_all).astype(np.int32)).to(dev)
            ref = splan.step(fin)
            torch.cuda.synchronize()
            errs = {"loss": abs(float(gl[0]) - float(ref["loss"][0])) / abs(float(ref["loss"][0]))}
            for k in wkeys:
                if k in ref and float(ref[k].abs().max()) > 0:
                    errs[k] = rel(wsum[k].reshape(-1), ref[k].reshape(-1))
            for k in ("d_t_pool", "d_xt_cls", "d_t_emb", "d_v_emb"):
                if k in ref and k in out:
                    errs[k] = rel(out[k], ref[k][:b])
            bad = {k: v for k, v in errs.items() if not v < 1e-3}
            print("dist check mode=%s graph=%s fusion=%s b=%d world=%d: %s" % (mode, use_graph, fusion, b, world, {k: "%.1e" % v for k, v in errs.items()}))
            if bad:
                print("FAIL", bad)
                ok = False
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
