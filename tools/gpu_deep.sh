#!/bin/bash
# deep-canopy case (BASELINE.json configs[4]: n_z = 1000): one bench line per scheme, 1184 scenarios x 2100 bands x 1000 levels
mkdir -p gpurun_out; : > gpurun_out/deep.txt
for sch in ${SCHEMES:-4s 2s zq n79 bf}; do
  timeout 600 python bench.py --scheme $sch --nz 1000 --scenarios 1184 --chunk 296 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/v.log 2>&1
  python - "$sch" <<'PY' | tee -a gpurun_out/deep.txt
import json, sys
l=[x for x in open("gpurun_out/v.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
    print("%-6s n_z=1000 value=%.3e frac=%.3f GB/s=%.0f kernel_ms=%.3f sm_mhz=%s kernel=%s" % (sys.argv[1], d["value"], r["frac"], r["achieved"], r["kernel_ms"], c.get("sm_mhz"), r.get("kernel")))
else:
    print(sys.argv[1], "FAILED"); print(open("gpurun_out/v.log").read()[-800:])
PY
done
