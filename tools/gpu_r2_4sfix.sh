#!/bin/bash
O=$PWD/gpurun_out/r2g; mkdir -p $O
show() { python - $1 $2 <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(sys.argv[2], "FAILED"); print(open(sys.argv[1].replace(".json",".err")).read()[-800:]); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]
print("%-22s value=%.4e frac=%.4f kernel_ms=%.3f sm=%s" % (sys.argv[2], d["value"], r["frac"], r["kernel_ms"], d["clocks"].get("sm_mhz")))
PY
}
timeout 1200 python -m pytest tests -m gpu -q -x -k "4s or flat or split or status or reduced or float32 or rows_kernel or deep" > $O/pytest_4s.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_4s.log
for rep in 1 2; do
timeout 600 python bench.py --scheme 4s --scenarios 132608 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > $O/fix_4s.json 2> $O/fix_4s.err; show $O/fix_4s.json 4s
done
timeout 600 python bench.py --scheme 4s --nz 1000 --scenarios 1184 --chunk 296 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > $O/dfix_4s.json 2> $O/dfix_4s.err; show $O/dfix_4s.json deep_4s
CMD="python bench.py --scheme 4s --scenarios 16576 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-legs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_4s_fix.csv $CMD > $O/ncu_4sfix.log 2>&1
grep -v "^==" $O/launches_4s_fix.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['Kernel Name'][:70], r['Grid Size'], r['Block Size'], r['Metric Value'], r['Metric Unit'])
" | tail -6
