#!/bin/bash
# full suite with the new defaults + bench lines + timeline for the profiles directory
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02h_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02h_pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for v in full itc fusion; do
  timeout 120 python scripts/prof_timeline.py --variant $v --out $O/r02h_timeline_c2_$v.txt > /dev/null 2> $O/r02h_tl.err || tail -3 $O/r02h_tl.err
done
head -3 $O/r02h_timeline_c2_full.txt | cut -c1-150
echo "--- A/B"
for env in "X=1" "TIC_PDL_CHAINS=0" "TIC_CODE_WARM=0" "TIC_PDL_CHAINS=0 TIC_CODE_WARM=0 TIC_ITC_FUSED_SMALL=0 TIC_CONCAT_PAIRWISE=0" "TIC_TIMELINE_SNAPSHOT=1"; do
  echo "$env"; env $env timeout 120 python scripts/timeline.py --replays 400 --plain-only --variant full,itc,fusion 2>&1 | grep "^workload"
done 2>&1 | tee $O/r02h_c2_ab.txt
timeout 600 python bench.py --steps 100 --warmup 5 > $O/r02h_bench_c2.json 2> $O/r02h_bench_c2.err; echo "bench c2 rc=$?"; tail -3 $O/r02h_bench_c2.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload c4 > $O/r02h_bench_c4.json 2> $O/r02h_bench_c4.err; echo "bench c4 rc=$?"; tail -3 $O/r02h_bench_c4.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02h_bench_ref_c2.json 2> /dev/null; echo "ref rc=$?"
python - <<'PY'
import json
for n in ("c2","c4","ref_c2"):
    try:
        d=json.loads(open("gpurun_out/r02h_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "ms/step %.4f value %.3e e2e %.3e launches %s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("config",{}).get("launches_per_step")))
        for k in d.get("kernels",[]): print("   %-70s %.4f ms  %.1f %s frac %.3f"%(k["kernel"][:70],k["ms"],k["achieved"],k["unit"],k["frac"]))
        for k in d.get("kernels_hbm_4096",[]):
            print("   %-70s %.4f ms  %.1f %s frac %.3f"%(k["kernel"][:70],k["ms"],k["achieved"],k["unit"],k["frac"]))
        if "e2e_dropin" in d: print("   dropin", d["e2e_dropin"]["ms_per_step"], d["e2e_dropin"]["value"])
        if "cpu_baseline" in d: print("   cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"])
    except Exception as e:
        print(n, "parse failed", e)
PY
