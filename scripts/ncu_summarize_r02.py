#!/usr/bin/env python
"""Summarise a multi-kernel `ncu --set full` report: one entry per captured launch, in capture order.

    python scripts/ncu_summarize_r02.py gpurun_out/r02_ncu_itc16k_d256.ncu-rep itc_fwd_16384x256 itc_bwd_16384x256 gemm_dv_16384x256 gemm_dt_16384x256

Writes profiles/r02_ncu_<key>.csv (the raw page of that launch), a readable profiles/r02_ncu_<report>.txt, and the keys in
profiles/ncu_traffic.json.  The tensor counter is sm__pipe_tensor_subpipe_hmma_cycles_active (the sub-pipe tcgen05.mma kind::f16 —
SASS UTCHMMA — executes on), not the `..._realtime ... elapsed` figure round 1 quoted."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
U = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
T = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
KEEP = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    rep, keys = sys.argv[1], sys.argv[2:]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    raw = "".join(l for l in raw.splitlines(True) if not l.startswith("=="))
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out_json = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    db = json.load(open(out_json)) if os.path.exists(out_json) else {}
    txt = ["ncu --set full --clock-control none, report %s" % os.path.basename(rep), ""]
    for key, vals in zip(keys, rows[2:]):
        d = dict(zip(hdr, zip(units, vals)))
        w = csv.writer(open(os.path.join(ROOT, "profiles", "r02_ncu_%s.csv" % key), "w"))
        w.writerows([hdr, units, vals])
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(d[h][1]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        g = lambda k: d.get(k, ("", ""))  # noqa: E731
        num = lambda k: float(g(k)[1].replace(",", "")) if g(k)[1] not in ("", "no data") else float("nan")  # noqa: E731
        rd, wr = num("dram__bytes_read.sum") * U.get(g("dram__bytes_read.sum")[0], 1), num("dram__bytes_write.sum") * U.get(g("dram__bytes_write.sum")[0], 1)
        ms = num("gpu__time_duration.sum") * T.get(g("gpu__time_duration.sum")[0], float("nan"))
        db[key] = {"kernel": g("Kernel Name")[1][:160], "launches": 1, "round": "r02", "ms_under_ncu": ms, "dram_read_bytes": rd,
                   "dram_write_bytes": wr, "traffic_bytes": rd + wr, "registers_per_thread": g("launch__registers_per_thread")[1],
                   "grid": g("launch__grid_size")[1], "block": g("launch__block_size")[1],
                   "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active")[1],
                   "tensor_hmma_subpipe_active_pct": g("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active")[1],
                   "xu_pipe_pct": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")[1],
                   "alu_pipe_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")[1],
                   "warp_stall_cycles_per_issue": ["%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:6]]}
        txt.append("== %s: %s" % (key, g("Kernel Name")[1][:110]))
        for k in KEEP:
            txt.append("   %-80s %s %s" % (k, g(k)[1], g(k)[0]))
        txt.append("   warp stall cycles per issued instruction: " + ", ".join(db[key]["warp_stall_cycles_per_issue"]))
        txt.append("")
    name = os.path.basename(rep).replace(".ncu-rep", "")
    open(os.path.join(ROOT, "profiles", "%s.txt" % name), "w").write("\n".join(txt))
    json.dump(db, open(out_json, "w"), indent=1, sort_keys=True)
    print("\n".join(txt))


if __name__ == "__main__":
    main()
