#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "4s or nonuniform or random" > $O/pytest_4s.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_4s.log
for sch in 4s 2s; do
timeout 600 python bench.py --scheme $sch --steps 3 --warmup 3 --no-cpu-baseline --host-sample 0 > $O/b_$sch.json 2> $O/b_$sch.err; echo "bench $sch rc=$?"
python - $O/b_$sch.json <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
d=json.loads(l[-1]); r=d["roofline"]
print("  value=%.4e ms/step=%.2f frac=%.3f kernel_ms=%.3f" % (d["value"], d["ms_per_step"], r["frac"], r["kernel_ms"]), {k:(v.get("value"), v.get("roofline_frac")) for k,v in d.items() if k in ("reduced_diagnostic","nonuniform_lai")})
PY
done
