#!/usr/bin/env python
"""Phase stamps of the fused ITC kernel INSIDE the replayed c2 step (tic_debug_set_trace), beside the same kernel alone."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import tic_b200.capi as capi  # noqa: E402
import tic_b200.plan as P  # noqa: E402
from timeline import capture  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
spec = bench.workload_spec("c2", 1)
host = bench.make_inputs(spec)
dev_in = {k: (v.to(torch.bfloat16) if k in bench.BF16_KEYS else v).to(dev) for k, v in host.items()}
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
for variant in ("itc", "itc+prewarm", "itc+prewarm_other_outputs"):
    prewarm = "prewarm" in variant
    other = variant.endswith("other_outputs")
    variant = variant.split("+")[0]
    fusion, use_itm = (spec["fusion"], True) if variant == "full" else (None, False)
    plan = P.HeadPlan(spec["B"], E=768, P=512, C=4, fusion=fusion, use_itc=True, use_itm=use_itm, Lv=1, device=dev)
    plan.bind_params({k: v.to(dev) for k, v in bench.synthetic_params(4, seed=40).items()}, live=True)
    for _ in range(3):
        plan.step(dev_in)
    rec = []
    orig = capi.call
    capi.call = P.call = lambda name, *a: (rec.append((name, a)), orig(name, *a))[1]
    plan.step(dev_in)
    capi.call = P.call = orig
    fused_call = [c for c in rec if c[0] == "tic_itc_fwd_bwd_small"][0]
    g = capture(plan, dev_in)
    scr = [torch.empty(256, 256, dtype=torch.bfloat16, device=dev) for _ in range(4)]
    tr = torch.zeros(16 * 8, dtype=torch.int64, device=dev)
    capi.call("tic_debug_set_trace", tr.data_ptr())
    rows = []
    for k in range(9):
        flush.fill_(float(k))
        torch.cuda.synchronize()
        if prewarm:    # the same kernel once, right after the flush: its code (and nothing else of the step) is back in L2
            a = list(fused_call[1]); a[-1] = torch.cuda.current_stream().cuda_stream
            if other:   # same code, but the four gradient operands go to scratch: is it the code or the output lines that were cold?
                n = len(a)
                for idx, t in ((n - 7, scr[0]), (n - 5, scr[1]), (n - 3, scr[2]), (n - 2, scr[3])):
                    a[idx] = t.data_ptr()
            orig(fused_call[0], *a)
            torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        t = tr.cpu().view(8, 16)
        t0 = int(t[:, 0].min())
        rows.append([max((int(t[c, i]) - t0) / 1e3 for c in range(8)) for i in (0, 1, 2, 8, 3, 4, 5, 6, 7)])
    capi.call("tic_debug_set_trace", None)
    med = [sorted(x)[len(x) // 2] for x in zip(*rows)]
    print("fused ITC kernel inside the replayed c2 step, variant %s%s (max over 8 CTAs, median of 9 replays), us since first CTA entry" % (variant, (" + the fused kernel run once after the flush" + (" (gradient operands to scratch)" if other else "")) if prewarm else ""))
    print("   entry  prolog  loads_issued  fwd_prefetch  acc_ready  fwd_done  cluster_sync  bwd_prefetch  end")
    print("   " + "  ".join("%6.2f" % x for x in med))
