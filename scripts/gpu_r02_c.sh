#!/bin/bash
# Round 2, GPU pass C: full suite + c2 timelines after the chain work + bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r02c_pytest_gpu.log | cut -c1-300
for v in full itc fusion; do
  timeout 120 python scripts/prof_timeline.py --variant $v --out $O/r02c_timeline_c2_$v.txt > /dev/null 2> $O/r02c_tl_$v.err || tail -3 $O/r02c_tl_$v.err
done
cut -c1-110 $O/r02c_timeline_c2_full.txt
echo "--- A/B (plain graph, 400 replays)"
for env in "X=1" "TIC_TIMELINE_SNAPSHOT=1" "TIC_PDL_CHAINS=1" "TIC_PDL=1"; do
  echo "$env"; env $env timeout 120 python scripts/timeline.py --replays 400 --plain-only --variant full,itc,fusion 2>&1 | grep "^workload"
done
timeout 600 python bench.py --steps 50 --warmup 5 > $O/r02c_bench_c2.json 2> $O/r02c_bench_c2.err; echo "bench c2 rc=$?"; tail -3 $O/r02c_bench_c2.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload c4 > $O/r02c_bench_c4.json 2> $O/r02c_bench_c4.err; echo "bench c4 rc=$?"; tail -3 $O/r02c_bench_c4.err
python - <<'PY'
import json
for n in ("c2","c4"):
    try:
        d=json.loads(open("gpurun_out/r02c_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "ms/step %.4f value %.3e e2e %.3e (sync %.4f) launches %s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["sync_ms_per_step"], d["config"]["launches_per_step"]))
        for k in d.get("kernels",[]): print("   %-70s %.4f ms  %.1f %s frac %.3f"%(k["kernel"][:70],k["ms"],k["achieved"],k["unit"],k["frac"]))
        for k in d.get("kernels_hbm_4096",[]):
            print("   %-70s %.4f ms  %.1f %s frac %.3f"%(k["kernel"][:70],k["ms"],k["achieved"],k["unit"],k["frac"]))
            if "pick_tiles_us" in k: print("      ", {a:round(b,1) for a,b in k.items() if a.endswith("_us")})
        if "e2e_dropin" in d: print("   dropin", d["e2e_dropin"]["ms_per_step"], d["e2e_dropin"]["value"])
    except Exception as e:
        print(n, "parse failed", e)
PY
