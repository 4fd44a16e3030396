#!/bin/bash
# session 5, call G: zq_pa two passes (closed-only kernel + Thomas pass on flagged columns): parity tests, A/B of resident-CTA targets
O=$PWD/gpurun_out/s5g; mkdir -p $O
line() { python - "$1" "$2" <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(sys.argv[2], "FAILED"); print(open(sys.argv[1]).read()[-600:]); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
print("%-34s value=%.4e frac=%.4f GB/s=%.0f kernel_ms=%.3f ms/step=%.2f chunk=%s sm_mhz=%s" % (sys.argv[2], d["value"], r["frac"], r["achieved"], r["kernel_ms"], d["ms_per_step"], d["config"].get("chunk"), c.get("sm_mhz")))
PY
}
timeout 900 python -m pytest tests -m gpu -q -x -k "zq_pa or two_pass or random or plugin or default or ragged or deep or float32" > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee $O/summary.txt; tail -3 $O/pytest.log
S="--scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
D="--nz 1000 --scenarios 1184 --chunk 296 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
run() { name=$1; sch=$2; args=$3; shift 3
  env "$@" timeout 300 python bench.py --scheme $sch $args > $O/v.json 2> $O/v.err; line $O/v.json "$name" | tee -a $O/summary.txt
}
for rep in 1 2; do
  run "zq_pa two-pass MINB3" zq_pa "$S" A=1
  run "zq_pa two-pass MINB2" zq_pa "$S" CRT1D_B200_LIB=$PWD/_r1/lib_zqpa_c2.so
  run "zq_pa two-pass MINB4" zq_pa "$S" CRT1D_B200_LIB=$PWD/_r1/lib_zqpa_c4.so
  run "zq_pa one-pass" zq_pa "$S" CRT1D_B200_ZQPA_ONE_PASS=1
done
run "deep_zq_pa two-pass MINB3" zq_pa "$D" A=1
run "deep_zq_pa two-pass MINB2" zq_pa "$D" CRT1D_B200_LIB=$PWD/_r1/lib_zqpa_c2.so
run "deep_zq_pa one-pass" zq_pa "$D" CRT1D_B200_ZQPA_ONE_PASS=1
CMD2="python bench.py --scheme zq_pa --scenarios 8288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"solve_kernel" -s 12 -c 2 -f -o /tmp/prof_zq_pa $CMD2 > $O/ncu_zq_pa.log 2>&1
echo "ncu zq_pa rc=$?" | tee -a $O/summary.txt
python tools/ncu_summary.py /tmp/prof_zq_pa.ncu-rep $O/ncu_full_zq_pa_two_pass.txt
