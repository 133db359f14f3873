#!/bin/bash
# final 1-GPU verification of the round: GPU suite, smoke, default bench (+ reference arm)
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02x_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02x_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02x_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02x_smoke.log
timeout 600 python bench.py > $O/r02x_bench_c2.json 2> $O/r02x_bench_c2.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --impl reference > $O/r02x_bench_ref_c2.json 2> $O/r02x_bench_ref.err; echo "bench ref rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02x_bench_c2.json").read().strip().splitlines()[-1])
print("c2 ms/step %.4f value %.3e e2e %.3e launches %s dropin %.3f ms"%(d["ms_per_step"],d["value"],d["e2e"]["value"],d["config"].get("launches_per_step"),d["e2e_dropin"]["ms_per_step"]))
for k in (d.get("kernels_hbm_4096") or []):
    print("   %-70s %.4f ms  %.1f %s frac %.3f traffic %s"%(k["kernel"][:70],k["ms"],k["achieved"],k["unit"],k["frac"],k.get("traffic")))
r=json.loads(open("gpurun_out/r02x_bench_ref_c2.json").read().strip().splitlines()[-1])
print("ref", r["value"], r["cpu_baseline"]["sample"][:100])
PY
