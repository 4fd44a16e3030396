#!/bin/bash
# A/B of the flat column mapping + checkpoint prefetch for the tridiagonal schemes; parity tests first.
O=$PWD/gpurun_out/r2g; mkdir -p $O
show() { python - $1 $2 <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(sys.argv[2], "FAILED"); print(open(sys.argv[1].replace(".json",".err")).read()[-800:]); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]
print("%-22s value=%.4e frac=%.4f kernel_ms=%.3f sm=%s %s" % (sys.argv[2], d["value"], r["frac"], r["kernel_ms"], d["clocks"].get("sm_mhz"), r.get("kernel")))
PY
}
timeout 900 python -m pytest tests -m gpu -q -x -k "flat or split or zq or n79 or deep or ragged" > $O/pytest_tri.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_tri.log
for rep in 1 2; do
for sch in zq n79 zq_pa; do
  timeout 300 python bench.py --scheme $sch --scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > $O/flat_$sch.json 2> $O/flat_$sch.err; show $O/flat_$sch.json flat_$sch
  CRT1D_B200_NO_FLAT=1 timeout 300 python bench.py --scheme $sch --scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > $O/tile_$sch.json 2> $O/tile_$sch.err; show $O/tile_$sch.json tile_$sch
done
done
for sch in zq zq_pa 4s; do
  timeout 600 python bench.py --scheme $sch --nz 1000 --scenarios 1184 --chunk 296 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > $O/dflat_$sch.json 2> $O/dflat_$sch.err; show $O/dflat_$sch.json deep_flat_$sch
done
