#!/usr/bin/env python
"""NVLink bytes per step of the multi-GPU plans, from the driver's per-link data counters (`nvidia-smi nvlink -gt d`) read
before and after S replays of the captured step — the counters ncu cannot give for a multi-process exchange.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 \
        scripts/nvlink_bytes.py --workload c2 --steps 2000 --out gpurun_out/r02_nvlink_c2_g2

Rank 0 writes <out>_before.txt / <out>_after.txt (raw nvidia-smi output, all GPUs) and <out>.json = {steps, world, algorithmic
bytes per rank and step}.  scripts/nvlink_bytes.py --parse <out> turns the pair into bytes per step and GPU."""
import argparse
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse(out):
    meta = json.load(open(out + ".json"))

    def totals(path):
        tx, rx, gpu = {}, {}, None
        for ln in open(path):
            m = re.match(r"\s*GPU (\d+):", ln)
            if m:
                gpu = int(m.group(1))
            m = re.search(r"Data Tx:\s*(\d+)\s*KiB", ln)
            if m and gpu is not None:
                tx[gpu] = tx.get(gpu, 0) + int(m.group(1)) * 1024
            m = re.search(r"Data Rx:\s*(\d+)\s*KiB", ln)
            if m and gpu is not None:
                rx[gpu] = rx.get(gpu, 0) + int(m.group(1)) * 1024
        return tx, rx
    tb, rb = totals(out + "_before.txt")
    ta, ra = totals(out + "_after.txt")
    res = dict(meta)
    res["per_gpu"] = {str(g): {"tx_bytes_per_step": (ta[g] - tb.get(g, 0)) / meta["steps"], "rx_bytes_per_step": (ra[g] - rb.get(g, 0)) / meta["steps"]}
                      for g in sorted(ta) if ta[g] != tb.get(g, 0) or ra[g] != rb.get(g, 0)}
    print(json.dumps(res, indent=1))
    json.dump(res, open(out + ".json", "w"), indent=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_nvlink"))
    ap.add_argument("--parse", default=None)
    args = ap.parse_args()
    if args.parse:
        return parse(args.parse)
    import torch
    import torch.distributed as dist
    import bench
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
    plan.bind_params({k: v.to(dev) for k, v in bench.synthetic_params(spec["C"], seed=40).items()}, live=True)
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
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    dist.barrier()

    def snap(tag):
        if rank == 0:
            r = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            open("%s_%s.txt" % (args.out, tag), "w").write(r.stdout)
        dist.barrier()
    snap("before")
    for _ in range(args.steps):
        g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    snap("after")
    if rank == 0:
        Pe, N = plan.Pe, B * world
        if plan.itc_mode == "symmetric":
            per_peer = 2 * B * Pe * 2 * (2 if plan.has_lo else 1) + 2 * B * 4 + 2 * B * 4     # embeddings (+ residuals), inverse norms, lse
            alg = {"sent_per_rank_and_step": per_peer * (world - 1), "what": "push form: rows (hi%s) + inverse norms + lse vectors to every remote peer" % ("+lo" if plan.has_lo else "")}
        else:
            alg = {"received_per_rank_and_step": (world - 1) * (B * Pe * 2 + B * 4 + N * 4 + B * Pe * 4),
                   "what": "row-block form: V rows + inverse norms + column sums [N] + the rank's slice of every peer's dV contributions [b,P] fp32"}
        json.dump({"workload": spec["workload"], "world": world, "steps": args.steps, "itc_mode": plan.itc_mode, "algorithmic": alg},
                  open(args.out + ".json", "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
