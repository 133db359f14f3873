#!/usr/bin/env python
"""BASELINE configs[4] (c5): the aux-loss (ITC) module sweep — GLOBAL batch 4k..64k x d 256..1024 — in ONE process per GPU
count (20 points each; the per-point bench.py launches of scripts/sweep_c5.sh cost a python start-up per point).

    python scripts/sweep_c5.py                                  # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29571 scripts/sweep_c5.py

The global batch B is split over the ranks (B/N rows each; strong scaling: the problem is the same 6*B^2*d FLOP at every N).
Every point: 3 warm-up + 6 timed replays of the captured step, L2 flushed between replays, CUDA events, max over ranks.
Output: gpurun_out/c5_sweep_g<N>.json (one record per point) and a table on stdout.  The CPU column (N = 1 run only) is the
oracle port on the host cores at B <= 8192, extrapolated with B^2 above (and labelled so), as BASELINE.md section 5 prescribes.
"""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=dev)
        dist = dist_mod
    import tic_b200.plan as P
    from tic_b200 import capi
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("bf16_tflops_sustained", 1418.0)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    scale = math.exp(2.6592)
    recs = []
    for d in (256, 512, 768, 1024):
        for B in (4096, 8192, 16384, 32768, 65536):
            b = B // world
            g = torch.Generator().manual_seed(40 + 1000 * rank)
            T = torch.randn(b, d, generator=g)
            inp = {"t_pool": T.to(torch.bfloat16).to(dev), "v_pool": (torch.randn(b, d, generator=g) + 0.25 * T).to(torch.bfloat16).to(dev)}
            if world == 1:
                plan = P.HeadPlan(b, E=768, P=None, C=4, fusion=None, use_itc=True, use_itm=False, Lv=1, device=dev)
                plan.itc = P.ItcPlan(b, b, d, dev)
                plan.itc.scale_dev = plan.scale_t
                plan.Pe = d
                plan.out["d_t_emb"], plan.out["d_v_emb"] = torch.empty(b, d, device=dev), torch.empty(b, d, device=dev)
            else:
                from tic_b200.peer import PeerHeadPlan
                plan = PeerHeadPlan(b, world=world, rank=rank, E=768, P=None, d=d, C=4, fusion=None, use_itc=True, use_itm=False, Lv=1,
                                    device=dev)
            plan.set_weights({"dual_encoder.logit_scale": torch.tensor(2.6592, device=dev)})
            for _ in range(3):
                plan.step(inp)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                plan.step(inp)
            torch.cuda.current_stream().wait_stream(s)
            with torch.cuda.graph(gr):
                plan.step(inp)
            gr.replay()
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(6)]
            if dist is not None:
                dist.barrier()
            for k in range(6):
                flush.fill_(float(k))
                if world > 1:
                    plan.pg.exchange("align")
                ev[k][0].record(); gr.replay(); ev[k][1].record()
            torch.cuda.synchronize()
            ms = sum(a.elapsed_time(c) for a, c in ev) / 6
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            loss = float(plan.global_loss()[2]) if world > 1 else float(plan.out["loss"][2])
            tf = 6.0 * B * B * d / (ms * 1e-3) * 1e-12
            rec = {"B_global": B, "d": d, "n_gpus": world, "rows_per_gpu": b, "ms_per_step": ms, "samples_per_s": B / (ms * 1e-3),
                   "algorithmic_tflops_total": tf, "frac_of_sustained_peak_per_gpu": tf / world / peak, "itc_loss": loss,
                   "mode": getattr(plan, "itc_mode", "single")}
            recs.append(rec)
            if rank == 0:
                print("B=%6d d=%5d N=%d  %9.3f ms  %10.3e samples/s  %8.1f TFLOP/s total  %.3f of sustained/GPU  loss %.4f (%s)" %
                      (B, d, world, ms, rec["samples_per_s"], tf, rec["frac_of_sustained_peak_per_gpu"], loss, rec["mode"]), flush=True)
            del gr
            if world > 1:
                torch.cuda.synchronize()
                dist.barrier()
                pg = plan.pg
                pg.close()
                capi.call("tic_peer_free", pg.local)
                dist.barrier()
            del plan, inp
            torch.cuda.empty_cache()
    if rank == 0 and world == 1:      # CPU column: oracle port at B <= 8192, all host cores
        from oracle import restatement as R
        torch.set_num_threads(os.cpu_count() or 1)
        cpu = {}
        for d in (256, 512, 768, 1024):
            for B in (4096, 8192):
                g = torch.Generator().manual_seed(40)
                T0, V0 = torch.randn(B, d, generator=g), torch.randn(B, d, generator=g)
                ls = torch.tensor(2.6592, requires_grad=True)
                ts = []
                for it in range(4):
                    t0 = time.perf_counter()
                    T, V = T0.clone().requires_grad_(True), V0.clone().requires_grad_(True)
                    R.clip_loss(R.itc_logits(T, V, ls)).backward()
                    ts.append(time.perf_counter() - t0)
                cpu[(B, d)] = sorted(ts[1:])[1]
        for r in recs:
            B, d = r["B_global"], r["d"]
            if (B, d) in cpu:
                r["cpu_samples_per_s"], r["cpu_note"] = B / cpu[(B, d)], "measured, %d cores" % (os.cpu_count() or 1)
            else:
                r["cpu_samples_per_s"] = B / (cpu[(8192, d)] * (B / 8192.0) ** 2)
                r["cpu_note"] = "extrapolated with B^2 from B=8192 (%d cores)" % (os.cpu_count() or 1)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(recs, open(os.path.join(ROOT, "gpurun_out", "c5_sweep_g%d.json" % world), "w"), indent=1)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
