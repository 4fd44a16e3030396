#!/bin/bash
# session 5, call H: does LOWER occupancy help the column-pattern kernels?  (CRT1D_B200_SMEM_PAD forces fewer resident CTAs)
O=$PWD/gpurun_out/s5h; mkdir -p $O
line() { python - "$1" "$2" <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(sys.argv[2], "FAILED"); print(open(sys.argv[1]).read()[-600:]); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
print("%-38s value=%.4e frac=%.4f GB/s=%.0f kernel_ms=%.3f ms/step=%.2f sm_mhz=%s" % (sys.argv[2], d["value"], r["frac"], r["achieved"], r["kernel_ms"], d["ms_per_step"], c.get("sm_mhz")))
PY
}
timeout 600 python -m pytest tests -m gpu -q -x -k "closed_form_and_thomas or preferred or flat_column" > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee $O/summary.txt; tail -3 $O/pytest.log
S="--scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
D="--nz 1000 --scenarios 1184 --chunk 296 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
run() { name=$1; sch=$2; args=$3; shift 3
  env "$@" timeout 300 python bench.py --scheme $sch $args > $O/v.json 2> $O/v.err; line $O/v.json "$name" | tee -a $O/summary.txt
}
for rep in 1 2; do
  run "zq 2x256 (default)" zq "$S" A=1
  run "zq 1x256" zq "$S" CRT1D_B200_SMEM_PAD=120000
  run "zq 2x128" zq "$S" CRT1D_B200_TILE_THREADS=128 CRT1D_B200_SMEM_PAD=60000
  run "n79 2x256 (default)" n79 "$S" A=1
  run "n79 1x256" n79 "$S" CRT1D_B200_SMEM_PAD=120000
  run "zq_pa 2x256 (default)" zq_pa "$S" A=1
  run "zq_pa 1x256" zq_pa "$S" CRT1D_B200_SMEM_PAD=120000
done
run "deep_zq 2x256 (default)" zq "$D" A=1
run "deep_zq 1x256" zq "$D" CRT1D_B200_SMEM_PAD=100000
