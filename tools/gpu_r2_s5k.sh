#!/bin/bash
# session 5, call K: deep zq with checkpoints every 20 levels (built in): parity + A/B against the 10-level spacing
O=$PWD/gpurun_out/s5k; mkdir -p $O
line() { python - "$1" "$2" <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(sys.argv[2], "FAILED"); print(open(sys.argv[1]).read()[-600:]); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
print("%-34s value=%.4e frac=%.4f GB/s=%.0f kernel_ms=%.3f ms/step=%.2f chunk=%s sm_mhz=%s" % (sys.argv[2], d["value"], r["frac"], r["achieved"], r["kernel_ms"], d["ms_per_step"], d["config"].get("chunk"), c.get("sm_mhz")))
PY
}
timeout 900 python -m pytest tests -m gpu -q -x -k "wide_checkpoint or deep or preferred or flat_column" > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee $O/summary.txt; tail -3 $O/pytest.log
run() { name=$1; sch=$2; args=$3; shift 3
  env "$@" timeout 300 python bench.py --scheme $sch $args > $O/v.json 2> $O/v.err; line $O/v.json "$name" | tee -a $O/summary.txt
}
for nz in 1000 600; do
  sc=1184; ch=296; if [ $nz = 600 ]; then sc=1776; ch=444; fi
  D="--nz $nz --scenarios $sc --chunk -$ch --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
  for rep in 1 2; do
    run "zq n_z=$nz CK=20 (default)" zq "$D" A=1
    run "zq n_z=$nz CK=10" zq "$D" CRT1D_B200_NO_WIDE_CK=1
  done
done
