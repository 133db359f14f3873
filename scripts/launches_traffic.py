#!/usr/bin/env python
"""Per-kernel DRAM traffic of the default (c2) workload from a cheap all-launch ncu pass:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/launches_c2_traffic.csv python bench.py --profile --no-graph --steps 2 --warmup 3
    python scripts/launches_traffic.py gpurun_out/launches_c2_traffic.csv 256x512

Takes the LAST step in the list and adds the keys bench.py looks up (itc_fwd_*, itc_bwd_*, gemm_dtdv_*, gemm_fusion_*) to
profiles/ncu_traffic.json."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNITS = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0}


def short(name):
    return name.split("(")[0][:120]


def is_gemm(name, a_mn, b_mn):
    """umma_gemm_kernel<BN, A_MN, B_MN, 4, StoreEpi, ...> with the given operand majors (ncu prints `64, 0, 1` or
    `(int)64, (bool)0, (bool)1` depending on the demangler)."""
    if "StoreEpi" not in name:
        return False
    args = name.split("<", 1)[1].replace("(int)", "").replace("(bool)", "").split(",")
    return int(args[1]) == a_mn and int(args[2]) == b_mn


def main():
    path, shape = sys.argv[1], sys.argv[2]
    lines = [l for l in open(path) if not l.startswith("==")]
    by_id = {}
    for r in csv.DictReader(lines):
        k = by_id.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"].replace(" ", "")})
        k[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNITS.get(r["Metric Unit"], 1.0)
    ids = sorted(by_id)
    last_fwd = max(i for i in ids if "ItcFwdEpi" in by_id[i]["name"])
    step = [i for i in ids if i >= last_fwd - 12]          # the launches around the last ITC forward = the last step
    out_json = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    db = json.load(open(out_json)) if os.path.exists(out_json) else {}

    def put(key, sel):
        if not sel:
            return
        rd = sum(by_id[i].get("dram__bytes_read.sum", 0.0) for i in sel)
        wr = sum(by_id[i].get("dram__bytes_write.sum", 0.0) for i in sel)
        ms = sum(by_id[i].get("gpu__time_duration.sum", 0.0) for i in sel)
        db[key] = {"kernel": short(by_id[sel[0]]["name"]), "launches": len(sel), "ms_under_ncu": ms, "dram_read_bytes": rd,
                   "dram_write_bytes": wr, "traffic_bytes": rd + wr, "source": os.path.basename(path)}
        print(key, db[key])

    put("itc_fwd_" + shape, [last_fwd])
    bwd = [i for i in step if "ItcBwdEpi" in by_id[i]["name"] and i > last_fwd]
    put("itc_bwd_" + shape, bwd[:1])
    if bwd:
        put("gemm_dtdv_" + shape, [i for i in ids if i > bwd[0] and is_gemm(by_id[i]["name"], 0, 1)][:2])
    put("gemm_fusion_" + shape, [i for i in step if is_gemm(by_id[i]["name"], 0, 0) and by_id[i]["grid"].startswith("(48,")][:1])
    json.dump(db, open(out_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
