#!/bin/bash
# throughput of kernel tuning variants (libcrt1d_b200_<name>.so built with different -D knobs)
mkdir -p gpurun_out
: > gpurun_out/variants.txt
for lib in "" mb4 mb5 b256mb2 b64mb8 b64mb10; do
  for sch in ${SCHEMES:-2s 4s}; do
    if [ -n "$lib" ]; then export CRT1D_B200_LIB=$PWD/crt1d_b200/libcrt1d_b200_$lib.so; else unset CRT1D_B200_LIB; fi
    timeout 300 python bench.py --scheme $sch --scenarios ${NSCEN:-131072} --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/v.log 2>&1
    python - "$lib" "$sch" <<'PY' | tee -a gpurun_out/variants.txt
import json, sys
l=[x for x in open("gpurun_out/v.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
    print("%-8s %-4s value=%.3e frac=%.3f GB/s=%.0f kernel_ms=%.3f sm_mhz=%s reasons=%s" % (sys.argv[1] or "default", sys.argv[2], d["value"], r["frac"], r["achieved"], r["kernel_ms"], c.get("sm_mhz"), c.get("reasons")))
else:
    print(sys.argv[1], sys.argv[2], "FAILED"); print(open("gpurun_out/v.log").read()[-800:])
PY
  done
done
