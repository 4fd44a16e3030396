#!/bin/bash
# A/B of two library builds on one box: $AB_LIBS = space-separated .so paths ("default" = the in-tree library);
# per scheme in $AB_SCHEMES one short bench line each, interleaved A B A B; optional reduced-diagnostic timing.
mkdir -p gpurun_out
: > gpurun_out/ab.txt
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/ab.txt; tail -3 gpurun_out/pytest_gpu.log
fi
for rep in 1 2; do
for sch in ${AB_SCHEMES:-2s 4s}; do
for lib in ${AB_LIBS:-default}; do
  if [ "$lib" = default ]; then unset CRT1D_B200_LIB; else export CRT1D_B200_LIB=$PWD/$lib; fi
  timeout 300 python bench.py --scheme $sch --scenarios ${AB_SCEN:-66304} --steps 3 --warmup 3 --no-cpu-baseline --no-e2e ${AB_ARGS} > gpurun_out/v.log 2>&1
  python - "$sch" "$lib" <<'PY' | tee -a gpurun_out/ab.txt
import json, sys
l=[x for x in open("gpurun_out/v.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
    print("%-6s %-40s value=%.4e frac=%.4f kernel_ms=%.3f sm_mhz=%s" % (sys.argv[1], sys.argv[2], d["value"], r["frac"], r["kernel_ms"], c.get("sm_mhz")))
else:
    print(sys.argv[1], sys.argv[2], "FAILED"); print(open("gpurun_out/v.log").read()[-800:])
PY
done; done; done
if [ -n "$AB_DIAG" ]; then
for lib in ${AB_LIBS:-default}; do
  if [ "$lib" = default ]; then unset CRT1D_B200_LIB; else export CRT1D_B200_LIB=$PWD/$lib; fi
  echo "== diag $lib" | tee -a gpurun_out/ab.txt
  timeout 300 python tools/diag_only.py $AB_DIAG 2>&1 | grep scheme | cut -c1-75 | tee -a gpurun_out/ab.txt
done
fi
