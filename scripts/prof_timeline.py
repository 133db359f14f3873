#!/usr/bin/env python
"""Kernel timeline of one step replayed as a CUDA graph, from CUPTI's concurrent-kernel activity records (torch.profiler):
start offset, duration and stream of every kernel of the LAST profiled replay — the critical chain of the small-batch step
without the event nodes scripts/timeline.py has to insert (those perturb the step by ~45 us).  A profiler run is never a
bench number: the un-profiled step time is printed beside it.

    python scripts/prof_timeline.py [--workload c2] [--variant full|itc|fusion] [--snapshot] [--replays 5] [--out file]
"""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import tic_b200.plan as P  # noqa: E402
from timeline import capture, time_graph  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--variant", default="full")
    ap.add_argument("--snapshot", action="store_true", help="set_weights (no in-step refresh) instead of live master weights")
    ap.add_argument("--replays", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    spec = bench.workload_spec(args.workload, 1)
    host = bench.make_inputs(spec)
    dev_in = {k: (v.to(torch.bfloat16) if k in bench.BF16_KEYS else v).to(dev) for k, v in host.items()}
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    use_itc, use_itm, fusion = spec["use_itc"], spec["use_itm"], spec["fusion"]
    if args.variant == "itc":
        fusion, use_itm = None, False
    if args.variant == "fusion":
        use_itc = False
    plan = P.HeadPlan(spec["B"], E=spec["E"], P=spec["P"], C=spec["C"], fusion=fusion, use_itc=use_itc, use_itm=use_itm,
                      Lv=max(spec["Lv"], 1), device=dev, itm_mode=spec.get("itm_mode", "uniform") if use_itc else "uniform")
    if spec["P"] is None:      # ITC-only sweep point ("itc:<B>x<d>"): the embeddings are the inputs
        B = spec["B"]
        plan.itc = P.ItcPlan(B, B, spec["d"], dev)
        plan.itc.scale_dev = plan.scale_t
        plan.Pe = spec["d"]
        plan.out["d_t_emb"], plan.out["d_v_emb"] = torch.empty(B, spec["d"], device=dev), torch.empty(B, spec["d"], device=dev)
    master = {k: v.to(dev) for k, v in bench.synthetic_params(spec["C"], seed=40).items()}
    if args.snapshot:
        plan.set_weights(master)
    else:
        plan.bind_params(master, live=True)
    for _ in range(3):
        plan.step(dev_in)
    torch.cuda.synchronize()
    g = capture(plan, dev_in)
    for _ in range(3):
        g.replay()
    plain_us = time_graph(g, flush, 200)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for k in range(args.replays):
            flush.fill_(float(k))
            torch.cuda.synchronize()
            g.replay()
            torch.cuda.synchronize()
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
    ev.sort(key=lambda e: e["ts"])
    # split into replays at the flush kernels (fill of 256 MiB: by far the longest elementwise kernel)
    fills = [i for i, e in enumerate(ev) if "FillFunctor" in e["name"] and e["dur"] > 20]
    last = ev[fills[-1] + 1:] if fills else ev
    t0 = min(e["ts"] for e in last)
    lines = ["workload %s variant %s (%s weights): un-profiled graph %.2f us/step; profiled replay spans %.2f us, %d kernels" %
             (args.workload, args.variant, "snapshot" if args.snapshot else "live", plain_us,
              max(e["ts"] + e["dur"] for e in last) - t0, len(last)),
             "%9s %8s %7s  %s" % ("start_us", "dur_us", "stream", "kernel")]
    for e in last:
        name = e["name"]
        for a, b in (("tic::umma_gemm_kernel", "umma"), ("tic::", ""), ("void ", "")):
            name = name.replace(a, b)
        lines.append("%9.2f %8.2f %7s  %s" % (e["ts"] - t0, e["dur"], e.get("args", {}).get("stream", "?"), name[:150]))
    txt = "\n".join(lines)
    print(txt)
    if args.out:
        open(args.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
