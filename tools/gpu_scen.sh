#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large_batch or full_size or sweep_sample or default_case_matches_reference" > gpurun_out/pytest_rows.log 2>&1; tail -25 gpurun_out/pytest_rows.log
: > gpurun_out/scen_variants.txt
run() {
  timeout 300 python bench.py --scheme 2s --scenarios ${NSCEN:-131072} --chunk ${CHUNK:-4144} --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/v.log 2>&1
  python - "$1" <<'PY' | tee -a gpurun_out/scen_variants.txt
import json, sys
l=[x for x in open("gpurun_out/v.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
    print("%-16s value=%.3e frac=%.3f GB/s=%.0f kernel_ms=%.3f sm_mhz=%s reasons=%s" % (sys.argv[1], d["value"], r["frac"], r["achieved"], r["kernel_ms"], c.get("sm_mhz"), c.get("reasons")))
else:
    print(sys.argv[1], "FAILED"); print(open("gpurun_out/v.log").read()[-600:])
PY
}
export CRT1D_B200_2S_KERNEL=tile; run tile
export CRT1D_B200_2S_KERNEL=rows
for cfg in 4,512,0 4,512,1 3,512,1 6,512,1 10,512,1 4,640,1 6,640,1 10,640,1 6,768,1 10,768,1 10,1024,1 4,384,1 6,384,1; do
  export CRT1D_B200_ROWS_CFG=$cfg; run "rows $cfg"
done
