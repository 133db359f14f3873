"""bench.py on CPU: the reference arm's JSON line (the contract keys the driver reads) and the workload table."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_line_has_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["higher_is_better"] is True
    assert line["metric"] and line["unit"] == "samples/s" and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"]["workload"].startswith("c2")                      # the default workload = BASELINE configs[1]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "warm-up" in cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None


def test_workload_table_names_what_baseline_names():
    import bench
    c2 = bench.workload_spec("c2", 1)
    assert (c2["B"], c2["E"], c2["P"], c2["fusion"], c2["use_itc"], c2["use_itm"]) == (256, 768, 512, "concat", True, True)
    c4 = bench.workload_spec("c4", 1)
    assert c4["fusion"] == "attention" and c4["B"] == 4096 and c4["Lv"] == 197 and c4.get("itm_mode") == "hard"
    assert "hard-negative" in c4["workload"]
    pt = bench.workload_spec("itc:16384x768", 1)
    assert pt["fusion"] is None and pt["P"] is None and (pt["B"], pt["d"]) == (16384, 768)
    # at N GPUs the per-GPU batch stays (weak scaling) and the label says so
    assert bench.workload_spec("c2", 8)["B"] == 256
