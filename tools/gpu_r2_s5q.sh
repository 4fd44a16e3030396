#!/bin/bash
# session 5, call Q: 4s fix-up kernel compiled for 14 / 20 (default) / 28 resident warps per SM
O=$PWD/gpurun_out/s5q; mkdir -p $O
line() { python - "$1" "$2" <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(sys.argv[2], "FAILED"); print(open(sys.argv[1]).read()[-600:]); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
print("%-34s value=%.4e frac=%.4f GB/s=%.0f kernel_ms=%.3f ms/step=%.2f sm_mhz=%s" % (sys.argv[2], d["value"], r["frac"], r["achieved"], r["kernel_ms"], d["ms_per_step"], c.get("sm_mhz")))
PY
}
: > $O/summary.txt
run() { name=$1; sch=$2; args=$3; shift 3
  env "$@" timeout 300 python bench.py --scheme $sch $args > $O/v.json 2> $O/v.err; line $O/v.json "$name" | tee -a $O/summary.txt
}
S="--scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
for rep in 1 2 3; do
  run "4s fix-up MINB=20 (default)" 4s "$S" A=1
  run "4s fix-up MINB=14" 4s "$S" CRT1D_B200_LIB=$PWD/_r1/lib_fix14.so
  run "4s fix-up MINB=28" 4s "$S" CRT1D_B200_LIB=$PWD/_r1/lib_fix28.so
done
