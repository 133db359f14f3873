#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02v_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02v_pytest_gpu.log
timeout 600 python bench.py > $O/r02v_bench_c2.json 2> $O/r02v_bench_c2.err; echo "bench c2 rc=$?"
python - <<'PY'
import json
for w in ("c2",):
    try:
        d=json.loads(open("gpurun_out/r02v_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w,"ms/step %.4f value %.3e e2e %.3e launches %s"%(d["ms_per_step"],d["value"],d["e2e"]["value"],d["config"].get("launches_per_step")))
        for k in (d.get("kernels_hbm_4096") or []):
            print("   %-70s %.4f ms  %.1f %s frac %.3f"%(k["kernel"][:70],k["ms"],k["achieved"],k["unit"],k["frac"]))
    except Exception as e: print(w,"failed",e)
PY
NCU="ncu --clock-control none --kernel-name-base demangled"
timeout 900 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k "regex:ce_bidir|gather_rows|itm_sample|itm_hard_locate|ItcPickEpi" -c 4000 --csv \
   --log-file $O/r02_hbm_launches.csv python bench.py --steps 2 --warmup 3 > $O/r02v_ncu_h.log 2>&1; echo "hbm traffic rc=$?"
wc -l $O/r02_hbm_launches.csv
