#!/bin/bash
# Round 2, GPU pass A (one B200): the whole -m gpu suite, smoke, then bench lines for c2 (default) and c4 (hard negatives).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r02a_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/r02a_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02a_smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > $O/r02a_bench_c2.json 2> $O/r02a_bench_c2.err; echo "bench c2 rc=$?"; tail -3 $O/r02a_bench_c2.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload c4 > $O/r02a_bench_c4.json 2> $O/r02a_bench_c4.err; echo "bench c4 rc=$?"; tail -3 $O/r02a_bench_c4.err
python - <<'PY'
import json
for n in ("c2","c4"):
    try:
        d=json.loads(open("gpurun_out/r02a_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "ms/step %.4f value %.3e e2e %.3e launches %s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["launches_per_step"]))
        for k in d.get("kernels",[]): print("   %-70s %.4f ms  %.1f %s frac %.3f"%(k["kernel"][:70],k["ms"],k["achieved"],k["unit"],k["frac"]))
        for k in d.get("kernels_hbm_4096",[]): print("   %-70s %.4f ms  %.1f %s frac %.3f"%(k["kernel"][:70],k["ms"],k["achieved"],k["unit"],k["frac"]))
        if "e2e_dropin" in d: print("   dropin", d["e2e_dropin"]["ms_per_step"], d["e2e_dropin"]["value"])
        print("   cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(n, "parse failed", e)
PY
