#!/usr/bin/env python
"""One eager step of a bench workload, for `ncu -k regex:... -c N python scripts/ncu_step.py --workload itc:16384x256`
(no warm-up, no graph: the profiler replays each captured kernel itself)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--steps", type=int, default=1)
    args = ap.parse_args()
    import tic_b200.plan as P
    dev = torch.device("cuda", 0)
    spec = bench.workload_spec(args.workload, 1)
    host = bench.make_inputs(spec)
    dev_in = {k: (v.to(torch.bfloat16) if k in bench.BF16_KEYS else v).to(dev) for k, v in host.items()}
    plan = P.HeadPlan(spec["B"], E=spec["E"], P=spec["P"], C=spec["C"], fusion=spec["fusion"], use_itc=spec["use_itc"],
                      use_itm=spec["use_itm"], Lv=max(spec["Lv"], 1), device=dev, itm_mode=spec.get("itm_mode", "uniform"))
    if spec["P"] is None:
        B = spec["B"]
        plan.itc = P.ItcPlan(B, B, spec["d"], dev)
        plan.itc.scale_dev = plan.scale_t
        plan.Pe = spec["d"]
        plan.out["d_t_emb"], plan.out["d_v_emb"] = torch.empty(B, spec["d"], device=dev), torch.empty(B, spec["d"], device=dev)
    plan.bind_params({k: v.to(dev) for k, v in bench.synthetic_params(spec["C"], seed=40).items()}, live=True)
    for _ in range(args.steps):
        plan.step(dev_in)
    torch.cuda.synchronize()
    print("loss", [float(x) for x in plan.out["loss"]])


if __name__ == "__main__":
    main()
