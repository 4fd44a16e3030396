#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
: > gpurun_out/rows_generic.txt
for sch in bl bf g77 4s; do
  for mode in tile rows rows384 rows256; do
    unset CRT1D_B200_NO_ROWS CRT1D_B200_ROWS_THREADS
    [ $mode = tile ] && export CRT1D_B200_NO_ROWS=1
    [ $mode = rows384 ] && export CRT1D_B200_ROWS_THREADS=384
    [ $mode = rows256 ] && export CRT1D_B200_ROWS_THREADS=256
    timeout 300 python bench.py --scheme $sch --scenarios 66304 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/v.log 2>&1
    python - "$sch" "$mode" <<'PY' | tee -a gpurun_out/rows_generic.txt
import json, sys
l=[x for x in open("gpurun_out/v.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
    print("%-4s %-8s value=%.3e frac=%.3f GB/s=%.0f kernel_ms=%.3f sm_mhz=%s" % (sys.argv[1], sys.argv[2], d["value"], r["frac"], r["achieved"], r["kernel_ms"], c.get("sm_mhz")))
else:
    print(sys.argv[1], sys.argv[2], "FAILED"); print(open("gpurun_out/v.log").read()[-500:])
PY
  done
done
