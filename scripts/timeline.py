#!/usr/bin/env python
"""Device-side timeline of one step replayed as a CUDA graph (there is no nsys in the image).

Every C-ABI call of the step is followed by an *external* CUDA event (torch.cuda.Event(external=True) becomes an event-record
node of the captured graph), plus one event at the start of the step.  After N replays the median offset of every event from
the step's start event is printed: `end` = when the kernel launched by that call had finished, `stream` = the branch it ran on.
The event nodes perturb the step a little (the un-instrumented step time is printed beside the instrumented one), so this is a
tool for finding the critical chain, never for bench numbers.

    python scripts/timeline.py [--workload c2] [--variant full|noitm|itc|fusion] [--replays 30]
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import tic_b200.capi as capi  # noqa: E402
import tic_b200.plan as P  # noqa: E402
import tic_b200.utils  # noqa: E402,F401


def build_plan(spec, dev, variant):
    use_itc, use_itm, fusion = spec["use_itc"], spec["use_itm"], spec["fusion"]
    if variant == "noitm":
        use_itm = False
    if variant == "itc":
        fusion, use_itm = None, False
    if variant == "fusion":
        use_itc = False
    plan = P.HeadPlan(spec["B"], E=spec["E"], P=spec["P"], C=spec["C"], fusion=fusion, use_itc=use_itc, use_itm=use_itm,
                      Lv=max(spec["Lv"], 1), device=dev)
    master = {k: v.to(dev) for k, v in bench.synthetic_params(spec["C"], seed=40).items()}
    if os.environ.get("TIC_TIMELINE_SNAPSHOT", "0") == "1":
        plan.set_weights(master)               # round-1 form: pre-cast weights, no refresh inside the step
    else:
        plan.bind_params(master, live=True)    # the step refreshes the bf16 working copies / exp(logit_scale) itself
    return plan


def capture(plan, dev_in):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        plan.step(dev_in)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        plan.step(dev_in)
    return g


def time_graph(g, flush, n):
    ts = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(n):
        flush.fill_(float(k))
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.mean(sorted(ts)[:max(1, len(ts) * 3 // 4)]) * 1e3     # mean of the fastest 3/4 (event ticks are ~1 us)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--variant", default="full", help="comma list of full|noitm|itc|fusion")
    ap.add_argument("--replays", type=int, default=30)
    ap.add_argument("--json", default=None)
    ap.add_argument("--plain-only", action="store_true", help="time the un-instrumented graph only")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    spec = bench.workload_spec(args.workload, 1)
    host = bench.make_inputs(spec)
    dev_in = {k: (v.to(torch.bfloat16) if k in bench.BF16_KEYS else v).to(dev) for k, v in host.items()}
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    out = []
    for variant in args.variant.split(","):
        try:
            out.append(run_variant(args, spec, dev, dev_in, flush, variant))
        except Exception as e:   # keep going: one unsupported variant must not lose the others
            print("variant %s failed: %r" % (variant, e))
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


def run_variant(args, spec, dev, dev_in, flush, variant):
    plan = build_plan(spec, dev, variant)
    for _ in range(3):
        plan.step(dev_in)
    torch.cuda.synchronize()
    g_plain = capture(plan, dev_in)
    for _ in range(3):
        g_plain.replay()
    plain_us = time_graph(g_plain, flush, args.replays)
    if args.plain_only:
        print("workload %s variant %s: plain graph %.2f us/step" % (args.workload, variant, plain_us))
        return {"workload": args.workload, "variant": variant, "plain_us": plain_us}

    # ---- instrumented capture
    marks = []      # (name, stream id, event)
    orig_call = capi.call

    def traced(name, *a):
        rc = orig_call(name, *a)
        s = torch.cuda.current_stream()
        e = torch.cuda.Event(enable_timing=True, external=True)
        e.record(s)
        marks.append((name, s.cuda_stream, e))
        return rc

    start = torch.cuda.Event(enable_timing=True, external=True)
    orig_body = plan._step_body

    def body(inp):
        start.record(torch.cuda.current_stream())
        return orig_body(inp)

    capi.call = P.call = traced
    plan._step_body = body
    marks.clear()
    g = capture(plan, dev_in)      # the eager pass inside capture() records too: keep the marks of the captured pass only
    n_calls = len(marks) // 2
    marks = marks[n_calls:]
    capi.call = P.call = orig_call
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    inst_us = time_graph(g, flush, 5)
    offs = [[] for _ in marks]
    for k in range(args.replays):
        flush.fill_(float(k))
        g.replay()
        torch.cuda.synchronize()
        for i, (_, _, e) in enumerate(marks):
            offs[i].append(start.elapsed_time(e) * 1e3)
    streams = {}
    rows = []
    for (name, sid, _), o in zip(marks, offs):
        b = streams.setdefault(sid, "s%d" % len(streams))
        rows.append({"call": name, "branch": b, "end_us": statistics.median(o)})
    # start of a kernel ~ end of the previous call on the same branch (or the step start): report the gap as `dur<=`
    last = {}
    for r in rows:
        r["since_prev_on_branch_us"] = r["end_us"] - last.get(r["branch"], 0.0)
        last[r["branch"]] = r["end_us"]
    print("workload %s variant %s: plain graph %.1f us/step, instrumented %.1f us/step, %d calls" %
          (args.workload, variant, plain_us, inst_us, len(rows)))
    for r in sorted(rows, key=lambda r: r["end_us"]):
        print("%8.1f us  (+%6.1f on %-3s)  %s" % (r["end_us"], r["since_prev_on_branch_us"], r["branch"], r["call"]))
    return {"workload": args.workload, "variant": variant, "plain_us": plain_us, "instrumented_us": inst_us, "calls": rows}


if __name__ == "__main__":
    main()
