#!/bin/bash
# 1 GPU: full GPU test suite, smoke, default bench + c4, launch list and DRAM-traffic pass of the side kernels
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02u_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02u_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02u_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02u_smoke.log
timeout 600 python bench.py > $O/r02u_bench_c2.json 2> $O/r02u_bench_c2.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --workload c4 > $O/r02u_bench_c4.json 2> $O/r02u_bench_c4.err; echo "bench c4 rc=$?"
python - <<'PY'
import json
for w in ("c2","c4"):
    try:
        d=json.loads(open("gpurun_out/r02u_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w,"ms/step %.4f value %.3e e2e %.3e launches %s"%(d["ms_per_step"],d["value"],d["e2e"]["value"],d["config"].get("launches_per_step")))
        for k in (d.get("kernels_hbm_4096") or []):
            print("   %-70s %.4f ms  %.1f %s frac %.3f"%(k["kernel"][:70],k["ms"],k["achieved"],k["unit"],k["frac"]))
    except Exception as e: print(w,"failed",e)
PY
NCU="ncu --clock-control none --kernel-name-base demangled"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/r02_launches_c2.csv python bench.py --steps 2 --warmup 3 > $O/r02u_ncu_l.log 2>&1; echo "launch list rc=$?"
timeout 900 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k "regex:ce_bidir|gather_rows|itm_sample|itm_hard_locate|ItcPickEpi" -c 400 --csv \
   --log-file $O/r02_hbm_launches.csv python bench.py --steps 2 --warmup 3 > $O/r02u_ncu_h.log 2>&1; echo "hbm traffic rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:ce_bidir_fwd_kernel -s 2 -c 1 -f -o $O/r02_ncu_ce_fwd_4096 python bench.py --steps 2 --warmup 3 > $O/r02u_ncu_f.log 2>&1; echo "ce full rc=$?"
wc -l $O/r02_launches_c2.csv $O/r02_hbm_launches.csv; ls -la $O/*.ncu-rep
