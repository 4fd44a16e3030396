#!/bin/bash
# session 5, call I: full gpu test tier + smoke on the binary with the black-soil pivot floor; one line for zq and zq_pa
O=$PWD/gpurun_out/s5i; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee $O/summary.txt; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt; tail -1 $O/smoke.log
for sch in zq zq_pa; do
  timeout 300 python bench.py --scheme $sch --scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > $O/v_$sch.json 2> $O/v.err
  python - $O/v_$sch.json $sch <<'PY' | tee -a $O/summary.txt
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
d=json.loads(l[-1]); r=d["roofline"]
print("%-8s value=%.4e frac=%.4f kernel_ms=%.3f" % (sys.argv[2], d["value"], r["frac"], r["kernel_ms"]))
PY
done
