#!/bin/bash
# session 5, call F: persistent flat kernel + closed-form zq (register diet): parity tests, A/B
O=$PWD/gpurun_out/s5f; mkdir -p $O
line() { python - "$1" "$2" <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(sys.argv[2], "FAILED"); print(open(sys.argv[1]).read()[-600:]); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
print("%-34s value=%.4e frac=%.4f GB/s=%.0f kernel_ms=%.3f ms/step=%.2f chunk=%s sm_mhz=%s" % (sys.argv[2], d["value"], r["frac"], r["achieved"], r["kernel_ms"], d["ms_per_step"], d["config"].get("chunk"), c.get("sm_mhz")))
PY
}
timeout 900 python -m pytest tests -m gpu -q -x -k "zq or n79 or random or flat or deep or nonuniform or ragged" > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee $O/summary.txt; tail -3 $O/pytest.log
S="--scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
D="--nz 1000 --scenarios 1184 --chunk -296 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
run() { # name scheme args env...
  name=$1; sch=$2; args=$3; shift 3
  env "$@" timeout 300 python bench.py --scheme $sch $args > $O/v.json 2> $O/v.err; line $O/v.json "$name" | tee -a $O/summary.txt
}
for rep in 1 2; do
  run "zq closed persist" zq "$S" A=1
  run "zq closed no-persist" zq "$S" CRT1D_B200_NO_PERSIST=1
  run "zq thomas persist" zq "$S" CRT1D_B200_LIB=$PWD/_r1/lib_thomas.so
  run "zq thomas no-persist" zq "$S" CRT1D_B200_LIB=$PWD/_r1/lib_thomas.so CRT1D_B200_NO_PERSIST=1
  run "zq closed persist 192thr" zq "$S" CRT1D_B200_LIB=$PWD/_r1/lib_zq192.so
  run "n79 persist" n79 "$S" A=1
  run "n79 no-persist" n79 "$S" CRT1D_B200_NO_PERSIST=1
done
run "deep_zq closed persist" zq "$D" A=1
run "deep_zq closed no-persist" zq "$D" CRT1D_B200_NO_PERSIST=1
run "deep_zq thomas persist" zq "$D" CRT1D_B200_LIB=$PWD/_r1/lib_thomas.so
run "deep_zq thomas no-persist" zq "$D" CRT1D_B200_LIB=$PWD/_r1/lib_thomas.so CRT1D_B200_NO_PERSIST=1
run "deep_zq closed persist 192thr" zq "$D" CRT1D_B200_LIB=$PWD/_r1/lib_zq192.so
run "deep_n79 persist" n79 "$D" A=1
run "deep_n79 no-persist" n79 "$D" CRT1D_B200_NO_PERSIST=1
