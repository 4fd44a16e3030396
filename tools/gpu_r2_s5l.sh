#!/bin/bash
# session 5, call L: full gpu tier + smoke on the final binary (deep zq wide checkpoint spacing built in) + deep-canopy lines
O=$PWD/gpurun_out/s5l; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee $O/summary.txt; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt; tail -1 $O/smoke.log
for sch in 4s 2s zq n79 zq_pa bf; do
  timeout 600 python bench.py --scheme $sch --nz 1000 --scenarios 1184 --chunk -296 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > $O/d_$sch.json 2> $O/v.err
  python - $O/d_$sch.json deep_$sch <<'PY' | tee -a $O/deep.txt
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(sys.argv[2], "FAILED"); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
print("%-10s value=%.4e frac=%.4f GB/s=%.0f kernel_ms=%.3f ms/step=%.2f sm_mhz=%s reasons=%s kernel=%s" % (sys.argv[2], d["value"], r["frac"], r["achieved"], r["kernel_ms"], d["ms_per_step"], c.get("sm_mhz"), c.get("reasons"), r.get("kernel")))
PY
done
