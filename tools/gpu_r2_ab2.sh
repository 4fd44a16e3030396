#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "shared_factorisation" > $O/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_multi.log
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 600 python bench.py --scenarios 132608 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-legs $BARGS > $O/ab_$name.json 2> $O/ab_$name.err
  python - $O/ab_$name.json $name <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(sys.argv[2], "FAILED", open(sys.argv[1].replace(".json",".err")).read()[-500:]); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]
print("%-22s value=%.4e frac=%.4f kernel_ms=%.3f sm=%s" % (sys.argv[2], d["value"], r["frac"], r["kernel_ms"], d["clocks"].get("sm_mhz")))
PY
}
for rep in 1 2; do
BARGS="--scheme 4s" run 4s_default A=1
BARGS="--scheme 4s" run 4s_v2_nogeneral CRT1D_B200_LIB=$PWD/crt1d_b200/libcrt1d_v2_nogeneral.so
BARGS="--scheme 4s" run 4s_v3_noentire CRT1D_B200_LIB=$PWD/crt1d_b200/libcrt1d_v3_noentire.so
BARGS="--scheme zq --order matrix" run zq_multi_grouped CRT1D_B200_MULTI=1
BARGS="--scheme zq --order matrix" run zq_single_grouped CRT1D_B200_MULTI=0
BARGS="--scheme zq --order spec" run zq_multi_specorder CRT1D_B200_MULTI=1
done
