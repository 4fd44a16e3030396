#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/phases.txt
for dbg in 0 1 2; do
  export CRT1D_B200_ROWS_DEBUG=$dbg
  timeout 300 python bench.py --scheme 2s --scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/v.log 2>&1
  python - "$dbg" <<'PY' | tee -a gpurun_out/phases.txt
import json, sys
l=[x for x in open("gpurun_out/v.log") if x.startswith("{")]
d=json.loads(l[-1]); r=d["roofline"]
print("dbg=%s kernel_ms=%.3f GB/s(alg)=%.0f sm_mhz=%s" % (sys.argv[1], r["kernel_ms"], r["achieved"], d["clocks"].get("sm_mhz")))
PY
done
