#!/usr/bin/env python
"""DRAM traffic of the HBM-bound side kernels (bench.py `kernels_hbm_4096`) from one metrics-only ncu pass:

    ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        -k regex:'ce_bidir|gather_rows|itm_sample|itm_hard_locate|ItcPickEpi' -c 400 --csv --log-file gpurun_out/r02_hbm_launches.csv \
        python bench.py --steps 2 --warmup 3
    python scripts/ncu_hbm_keys.py gpurun_out/r02_hbm_launches.csv 4096

Adds ce_fwd_<B>, ce_bwd_<B>, itm_sample_gather_<B>, gather_rows_<B>, hard_locate_<B>, hard_pick_<B> to profiles/ncu_traffic.json
(per launch: the median over the captured launches of that kernel; ce_fwd sums its three kernels)."""
import csv
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNITS = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0}


def main():
    path, B = sys.argv[1], int(sys.argv[2])
    lines = [l for l in open(path) if not l.startswith("==")]
    by_id = {}
    for r in csv.DictReader(lines):
        k = by_id.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"].replace(" ", "")})
        k[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNITS.get(r["Metric Unit"], 1.0)
    out_json = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    db = json.load(open(out_json)) if os.path.exists(out_json) else {}

    def med(sel, m):
        return statistics.median(by_id[i].get(m, 0.0) for i in sel) if sel else 0.0

    def grid_of(i):
        return [int(x) for x in by_id[i]["grid"].strip("()").split(",")]

    def put(key, patterns, grid_min=0, grid_y=None):
        rd = wr = ms = 0.0
        names, n = [], 0
        for pat in patterns:
            sel = [i for i in by_id if pat in by_id[i]["name"] and grid_of(i)[0] >= grid_min and (grid_y is None or grid_of(i)[1] == grid_y)]
            if not sel:
                continue
            n = max(n, len(sel))
            rd += med(sel, "dram__bytes_read.sum"); wr += med(sel, "dram__bytes_write.sum"); ms += med(sel, "gpu__time_duration.sum")
            names.append(by_id[sel[0]]["name"].split("(")[0][:80])
        if not names:
            return
        db[key] = {"kernel": " + ".join(names), "launches": len(patterns), "captured": n, "ms_under_ncu": ms, "dram_read_bytes": rd,
                   "dram_write_bytes": wr, "traffic_bytes": rd + wr, "source": os.path.basename(path)}
        print(key, db[key])

    # only the launches of the B-row leg (the drop-in leg of the same bench run launches the same kernels at B = 256)
    put("ce_fwd_%d" % B, ["ce_bidir_fwd_kernel"], grid_min=B // 128)
    fw = db.get("ce_fwd_%d" % B)
    put("ce_lse_%d" % B, ["ce_bidir_lse_kernel"], grid_min=B // 256)
    ls = db.pop("ce_lse_%d" % B, None)
    if fw and ls:      # tic_ce_bidir_fwd = statistics kernel + combine kernel
        for k in ("ms_under_ncu", "dram_read_bytes", "dram_write_bytes", "traffic_bytes"):
            fw[k] += ls[k]
        fw["kernel"] += " + " + ls["kernel"]
        fw["launches"] = 2
    put("ce_bwd_%d" % B, ["ce_bidir_bwd_kernel"], grid_min=2, grid_y=B)
    put("itm_sample_gather_%d" % B, ["gather_rows_kernel"], grid_min=B // 8, grid_y=2)    # sampler fused into the pair gather
    put("gather_rows_%d" % B, ["gather_rows_kernel"], grid_min=B // 8, grid_y=1)
    put("hard_locate_%d" % B, ["itm_hard_locate_kernel"], grid_min=B // 32)
    put("hard_pick_%d" % B, ["ItcPickEpi"], grid_min=64)
    json.dump(db, open(out_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
