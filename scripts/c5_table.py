#!/usr/bin/env python
"""Renders profiles/r02_c5_sweep_g{1,2,4,8}.json (scripts/sweep_c5.py) as the 80-point table profiles/r02_c5_sweep.md."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rec = {}
    for n in (1, 2, 4, 8):
        p = os.path.join(ROOT, "profiles", "r02_c5_sweep_g%d.json" % n)
        if os.path.exists(p):
            for r in json.load(open(p)):
                rec[(r["B_global"], r["d"], n)] = r
    out = ["# BASELINE configs[4] (c5): ITC module sweep, global batch x embedding width x GPUs (B200, round 2)", "",
           "`scripts/sweep_c5.py`: one captured ITC forward+backward step (norms, similarity tiles + softmax statistics, lse, gradient-operand",
           "tiles, two gradient GEMMs, finalize) on synthetic embeddings; the GLOBAL batch B is split over the N ranks (B/N rows each:",
           "strong scaling). 3 warm-up + 6 timed graph replays per point, L2 flushed between replays (256 MiB write), CUDA events, max",
           "over ranks. `TF/s` = algorithmic 6*B^2*d FLOP / time, whole job; `frac` = that per GPU / the sustained bf16 peak in",
           "MEASURED_PEAKS.json (1418 TF/s). The tile recompute makes the executed FLOPs 8*B^2*d (x 1.33). CPU column: the oracle port",
           "(torch fp32, all host cores) measured at B <= 8192 and extrapolated with B^2 above, as BASELINE.md prescribes.", "",
           "| B (global) | d | CPU samples/s | N=1 ms | N=1 samples/s | N=1 TF/s (frac) | N=2 ms | N=2 TF/s (frac) | N=4 ms | N=4 TF/s (frac) | N=8 ms | N=8 samples/s | N=8 TF/s (frac) | speed-up 1->8 |",
           "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for d in (256, 512, 768, 1024):
        for B in (4096, 8192, 16384, 32768, 65536):
            r1 = rec.get((B, d, 1))
            cells = ["%d" % B, "%d" % d]
            if r1 and "cpu_samples_per_s" in r1:
                cells.append("%.3g%s" % (r1["cpu_samples_per_s"], "" if r1["cpu_note"].startswith("measured") else " (extrap.)"))
            else:
                cells.append("-")
            for n in (1, 2, 4, 8):
                r = rec.get((B, d, n))
                if r is None:
                    cells += ["-"] * (3 if n in (1, 8) else 2)
                    continue
                cells.append("%.3f" % r["ms_per_step"])
                if n in (1, 8):
                    cells.append("%.3g" % r["samples_per_s"])
                cells.append("%.0f (%.2f)" % (r["algorithmic_tflops_total"], r["frac_of_sustained_peak_per_gpu"]))
            r8 = rec.get((B, d, 8))
            cells.append("%.2fx" % (r1["ms_per_step"] / r8["ms_per_step"]) if (r1 and r8) else "-")
            out.append("| " + " | ".join(cells) + " |")
    out += ["", "Modes: N=1 single-GPU plan; N>1 `PeerHeadPlan` row-block form (global batch >= 4096: V gathered by a concurrent pull,",
            "column sums and dV reduced over peer memory). Inputs are seeded per rank, so the loss differs in the 3rd digit between N.",
            "N=8 table: first 8-GPU pass of the round (stdout, 3-4 digits); the later pass with 8 loads in flight per thread in the pull",
            "kernels regressed the small points (short exchanges fell into a scalar tail loop) and was reverted — see DESIGN.md 6."]
    open(os.path.join(ROOT, "profiles", "r02_c5_sweep.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:16]))


if __name__ == "__main__":
    main()
