#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/tilecfg.txt
export CRT1D_B200_2S_KERNEL=tile
for sch in ${SCHEMES:-2s 4s bl bf g77 zq n79 zq_pa}; do
  for cfg in 128,1 128,4 256,2; do
    export CRT1D_B200_TILE_CFG=$cfg
    timeout 300 python bench.py --scheme $sch --scenarios ${NSCEN:-65536} --chunk 4096 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/v.log 2>&1
    python - "$sch" "$cfg" <<'PY' | tee -a gpurun_out/tilecfg.txt
import json, sys
l=[x for x in open("gpurun_out/v.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
    print("%-6s %-6s value=%.3e frac=%.3f GB/s=%.0f kernel_ms=%.3f sm_mhz=%s" % (sys.argv[1], sys.argv[2], d["value"], r["frac"], r["achieved"], r["kernel_ms"], c.get("sm_mhz")))
else:
    print(sys.argv[1], sys.argv[2], "FAILED"); print(open("gpurun_out/v.log").read()[-500:])
PY
  done
done
