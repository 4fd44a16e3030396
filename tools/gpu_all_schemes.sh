#!/bin/bash
# GPU test tier + one bench line per scheme (66 304 scenarios) + ncu --set full of the tridiagonal tile kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/all_schemes.txt
for sch in 2s bl 4s bf g77 zq n79 zq_pa; do
  timeout 300 python bench.py --scheme $sch --scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/v.log 2>&1
  python - "$sch" <<'PY' | tee -a gpurun_out/all_schemes.txt
import json, sys
l=[x for x in open("gpurun_out/v.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
    print("%-6s value=%.3e frac=%.3f GB/s=%.0f kernel_ms=%.3f sm_mhz=%s reasons=%s" % (sys.argv[1], d["value"], r["frac"], r["achieved"], r["kernel_ms"], c.get("sm_mhz"), c.get("reasons")))
else:
    print(sys.argv[1], "FAILED"); print(open("gpurun_out/v.log").read()[-500:])
PY
done
if [ -z "$SKIP_NCU" ]; then
for sch in ${NCU_SCHEMES:-zq n79 zq_pa}; do
  CMD="python bench.py --scheme $sch --scenarios 8288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
  timeout 600 $CMD > gpurun_out/plain_$sch.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"solve_.*kernel" -s 6 -c 1 -f -o gpurun_out/prof_$sch $CMD > gpurun_out/ncu_$sch.log 2>&1
  echo "ncu $sch rc=$?"
done
fi
