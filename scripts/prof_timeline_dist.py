#!/usr/bin/env python
"""Kernel timeline (CUPTI via torch.profiler) of the multi-GPU c2 / c3 step on rank 0, launched under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \\
        scripts/prof_timeline_dist.py --workload c2 --out gpurun_out/timeline_c2_g2.txt
"""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--out", default=None)
    ap.add_argument("--replays", type=int, default=6)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from tic_b200.peer import PeerHeadPlan
    spec = bench.workload_spec(args.workload, world)
    host = bench.make_inputs(spec, rank=rank)
    dev_in = {k: (v.to(torch.bfloat16) if k in bench.BF16_KEYS else v).to(dev) for k, v in host.items()}
    B = spec["B"]
    plan = PeerHeadPlan(B, world=world, rank=rank, E=spec["E"], P=spec["P"] if spec["P"] is not None else None, d=spec.get("d"),
                        C=spec["C"], fusion=spec["fusion"], use_itc=spec["use_itc"], use_itm=spec["use_itm"], Lv=spec["Lv"], device=dev)
    master = {k: v.to(dev) for k, v in bench.synthetic_params(spec["C"], seed=40).items()}
    if os.environ.get("TIC_BENCH_SNAPSHOT", "0") == "1":
        plan.set_weights(master)
    else:
        plan.bind_params(master, live=True)      # as bench.py: fp32 masters refreshed by the root launch of the step
    for _ in range(3):
        plan.step(dev_in)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        plan.step(dev_in)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        plan.step(dev_in)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for k in range(100):
        flush.fill_(float(k))
        plan.pg.exchange("align")
        e0.record(); g.replay(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    plain = sorted(ts)[len(ts) // 2] * 1e3
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for k in range(args.replays):
            flush.fill_(float(k))
            plan.pg.exchange("align")
            torch.cuda.synchronize()
            g.replay()
            torch.cuda.synchronize()
    if rank == 0:
        path = os.path.join(tempfile.mkdtemp(), "trace.json")
        prof.export_chrome_trace(path)
        ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
        ev.sort(key=lambda e: e["ts"])
        fills = [i for i, e in enumerate(ev) if "FillFunctor" in e["name"] and e["dur"] > 20]
        last = ev[fills[-1] + 2:] if fills else ev      # + the align exchange
        t0 = min(e["ts"] for e in last)
        lines = ["workload %s on %d GPUs (rank 0): un-profiled graph median %.2f us/step; profiled replay spans %.2f us, %d kernels" %
                 (args.workload, world, plain, max(e["ts"] + e["dur"] for e in last) - t0, len(last)),
                 "%9s %8s %7s  %s" % ("start_us", "dur_us", "stream", "kernel")]
        for e in last:
            name = e["name"]
            for a, b in (("tic::umma_gemm_kernel", "umma"), ("tic::", ""), ("void ", "")):
                name = name.replace(a, b)
            lines.append("%9.2f %8.2f %7s  %s" % (e["ts"] - t0, e["dur"], e.get("args", {}).get("stream", "?"), name[:120]))
        txt = "\n".join(lines)
        print(txt)
        if args.out:
            open(args.out, "w").write(txt + "\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
