#!/bin/bash
# session 5, call M: deep 4s (BASELINE configs[4]): band split 1 (one 512-thread CTA per SM, whole rows) vs the default 3-way split
O=$PWD/gpurun_out/s5m; mkdir -p $O
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
D="--nz 1000 --scenarios 1184 --chunk 296 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
for rep in 1 2; do
  run "deep_4s default (3-way split)" 4s "$D" A=1
  run "deep_4s split=1" 4s "$D" CRT1D_B200_4S_SPLIT=1
  run "deep_4s split=4" 4s "$D" CRT1D_B200_4S_SPLIT=4
  run "deep_4s fixup parts=32" 4s "$D" CRT1D_B200_FIXUP_PARTS=32
  run "deep_4s fixup parts=8" 4s "$D" CRT1D_B200_FIXUP_PARTS=8
done
