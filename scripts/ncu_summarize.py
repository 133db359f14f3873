#!/usr/bin/env python
"""Summarises `ncu --set full` reports (.ncu-rep) into small tracked files under profiles/:

    python scripts/ncu_summarize.py <round-tag> gpurun_out/prof_*.ncu-rep

For every report: profiles/<tag>_ncu_<name>.csv  = the raw-page CSV (one row per captured launch, all metrics), and one entry
in profiles/ncu_traffic.json = {kernel key: {ms, dram_read_bytes, dram_write_bytes, traffic_bytes, issue_active_pct,
top_stalls, registers, grid}} which bench.py reads to fill `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum
of ONE launch of that kernel at the named workload; never a timing source — times under ncu are cold-cache and serialised)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def unit_bytes(unit, val):
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit)
    return None if mul is None else float(val.replace(",", "")) * mul


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    out_json = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    db = json.load(open(out_json)) if os.path.exists(out_json) else {}
    for rep in reps:
        name = os.path.basename(rep).replace(".ncu-rep", "").replace("prof_", "")
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
        raw = "".join(l for l in raw.splitlines(True) if not l.startswith("=="))
        open(os.path.join(ROOT, "profiles", "%s_ncu_%s.csv" % (tag, name)), "w").write(raw)
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = dict(zip(hdr, zip(units, vals)))
        g = lambda k: d.get(k, ("", ""))  # noqa: E731
        extra = [dict(zip(hdr, zip(units, r))) for r in rows[3:] if len(r) == len(hdr)]   # further captured launches
        stalls = []
        for h in hdr:
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    stalls.append((float(d[h][1].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        rd, wr = unit_bytes(*g("dram__bytes_read.sum")), unit_bytes(*g("dram__bytes_write.sum"))
        dur_unit, dur = g("gpu__time_duration.sum")
        ms = float(dur.replace(",", "")) * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(dur_unit, float("nan"))
        for e in extra:   # a key that covers several launches (e.g. the two gradient GEMMs): sum them
            rd += unit_bytes(*e["dram__bytes_read.sum"]) or 0
            wr += unit_bytes(*e["dram__bytes_write.sum"]) or 0
            u2, v2 = e["gpu__time_duration.sum"]
            ms += float(v2.replace(",", "")) * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(u2, float("nan"))
        db[name] = {
            "kernel": g("Kernel Name")[1][:160], "launches": 1 + len(extra), "round": tag, "ms_under_ncu": ms,
            "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": (rd or 0) + (wr or 0), "registers_per_thread": g("launch__registers_per_thread")[1],
            "grid": g("launch__grid_size")[1], "block": g("launch__block_size")[1],
            "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active")[1],
            "tensor_pipe_pct": g("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed")[1],
            "dram_throughput_pct": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")[1],
            "top_stalls": ["%s %.1f%%" % (n, 100 * v / tot) for v, n in sorted(stalls, reverse=True)[:5]],
        }
        print(name, json.dumps(db[name])[:400])
    json.dump(db, open(out_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
