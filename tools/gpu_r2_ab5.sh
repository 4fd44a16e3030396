#!/bin/bash
O=$PWD/gpurun_out/r2; mkdir -p $O
show() { python - $1 $2 <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(sys.argv[2], "FAILED"); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]
print("%-22s value=%.4e frac=%.4f kernel_ms=%.3f sm=%s" % (sys.argv[2], d["value"], r["frac"], r["kernel_ms"], d["clocks"].get("sm_mhz")))
PY
}
for rep in 1 2; do
(cd _r1 && timeout 600 python bench.py --scheme 4s --scenarios 132608 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/ab_r1_4s.json 2> $O/ab_r1_4s.err); show $O/ab_r1_4s.json r1_4s
timeout 600 python bench.py --scheme 4s --scenarios 132608 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > $O/ab_now_4s.json 2> $O/ab_now_4s.err; show $O/ab_now_4s.json now_4s
done
tail -3 $O/ab_r1_4s.err
