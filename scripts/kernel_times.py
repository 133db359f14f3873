#!/usr/bin/env python
"""Every C-ABI call of the c2 step timed ALONE, cold L2 (graph of flush+call minus graph of flush) and warm (graph of calls
back to back), plus the phase stamps of the fused ITC kernel.  Tells a slow kernel from a kernel that waits.

    python scripts/kernel_times.py [--workload c2]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import tic_b200.capi as capi  # noqa: E402
import tic_b200.plan as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    spec = bench.workload_spec(args.workload, 1)
    host = bench.make_inputs(spec)
    dev_in = {k: (v.to(torch.bfloat16) if k in bench.BF16_KEYS else v).to(dev) for k, v in host.items()}
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    plan = P.HeadPlan(spec["B"], E=spec["E"], P=spec["P"], C=spec["C"], fusion=spec["fusion"], use_itc=spec["use_itc"],
                      use_itm=spec["use_itm"], Lv=max(spec["Lv"], 1), device=dev, itm_mode=spec.get("itm_mode", "uniform"))
    master = {k: v.to(dev) for k, v in bench.synthetic_params(spec["C"], seed=40).items()}
    plan.bind_params(master, live=True)
    plan.parallel_streams = False          # program order on one stream: every call sees its inputs ready
    calls = []
    orig = capi.call

    def rec(name, *a):
        calls.append((name, a))
        return orig(name, *a)

    for _ in range(2):
        plan.step(dev_in)
    torch.cuda.synchronize()
    capi.call = P.call = rec
    plan.step(dev_in)
    capi.call = P.call = orig
    torch.cuda.synchronize()
    st = torch.cuda.current_stream().cuda_stream
    print("%-28s %9s %9s" % ("call (alone, one stream)", "cold_us", "warm_us"))
    tot_c = tot_w = 0.0
    for name, a in calls:
        if name in ("tic_gemm_plan",):
            continue
        a = list(a)
        a[-1] = None      # stream argument: replaced below by the capturing stream

        def fn(name=name, a=a):
            a2 = list(a)
            a2[-1] = torch.cuda.current_stream().cuda_stream
            orig(name, *a2)
        cold = bench.time_kernel_graph(fn, flush, reps=6, rounds=4) * 1e3
        # warm: 20 calls back to back in one graph
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            for _ in range(20):
                fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(4):
            e0.record(); g.replay(); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        warm = best / 20 * 1e3
        tot_c += cold; tot_w += warm
        extra = ""
        if name == "tic_gemm_bf16" or name == "tic_gemm_bf16_rowss":
            extra = "M=%d N=%d K=%d lo=%s/%s" % (a[12], a[13], a[14], bool(a[1]), bool(a[5]))
        print("%-28s %9.2f %9.2f  %s" % (name, cold, warm, extra))
    print("%-28s %9.2f %9.2f" % ("sum", tot_c, tot_w))
    # ---- phase stamps of the fused ITC kernel
    if getattr(plan, "_itc_bwd_fused", False):
        tr = torch.zeros(16 * 8, dtype=torch.int64, device=dev)
        capi.call("tic_debug_set_trace", tr.data_ptr())
        name, a = [c for c in calls if c[0] == "tic_itc_fwd_bwd_small"][0]
        a2 = list(a); a2[-1] = st
        others = [c for c in calls if c[0] in ("tic_heads_fwd_bwd", "tic_gemm_bf16", "tic_fusion_pair_grad", "tic_itc_grad_finalize")]

        def thrash():          # other kernels of the step on every SM: evicts this kernel's code from the SM-level caches only
            for nm, aa in others:
                aa2 = list(aa); aa2[-1] = st
                orig(nm, *aa2)
        modes = (("cold L2 left DIRTY by a 256 MiB write (the bench's flush)", lambda: flush.fill_(1.0)),
                 ("cold L2 left CLEAN by a 256 MiB read", lambda: flush.sum()),
                 ("warm L2, SM instruction caches thrashed by the step's other kernels", thrash),
                 ("warm (same kernel back to back)", lambda: orig(name, *a2)))
        for label, prep in modes:
            med = []
            for rep in range(5):
                prep()
                torch.cuda.synchronize()
                orig(name, *a2)
                torch.cuda.synchronize()
                t = tr.cpu().view(8, 16)
                t0 = int(t[:, 0].min())
                med.append([max((int(t[c, i]) - t0) / 1e3 for c in range(8)) for i in (0, 1, 2, 8, 3, 4, 5, 6, 7)])
            med = [sorted(x)[len(x) // 2] for x in zip(*med)]
            print("fused ITC kernel phases, %s: us since the first CTA's entry (max over the 8 CTAs, median of 5)" % label)
            print("   entry  prolog  loads_issued  fwd_prefetch  acc_ready  fwd_done  cluster_sync  bwd_prefetch  end")
            print("   " + "  ".join("%6.2f" % x for x in med))
        capi.call("tic_debug_set_trace", None)


if __name__ == "__main__":
    main()
