#!/bin/bash
# session 5, call B: deep zq with wider checkpoint spacing (one resident CTA, smaller parked-checkpoint working set);
# ncu --set full of the flat n79 / zq kernels and the zq_pa tile kernel (60 levels) with source pages
O=$PWD/gpurun_out/s5b; mkdir -p $O
line() { python - "$1" "$2" <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(sys.argv[2], "FAILED"); print(open(sys.argv[1]).read()[-600:]); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
print("%-28s value=%.4e frac=%.4f GB/s=%.0f kernel_ms=%.3f ms/step=%.2f chunk=%s sm_mhz=%s" % (sys.argv[2], d["value"], r["frac"], r["achieved"], r["kernel_ms"], d["ms_per_step"], d["config"].get("chunk"), c.get("sm_mhz")))
PY
}
: > $O/summary.txt
D="--nz 1000 --scenarios 1184 --chunk 296 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
for rep in 1 2; do
for lib in default lib_ck14 lib_ck16 lib_ck20; do
  if [ "$lib" = default ]; then unset CRT1D_B200_LIB; else export CRT1D_B200_LIB=$PWD/_r1/$lib.so; fi
  timeout 300 python bench.py --scheme zq $D > $O/v.json 2> $O/v.err; line $O/v.json "deep_zq $lib" | tee -a $O/summary.txt
done; done
unset CRT1D_B200_LIB
for sch in n79 zq_pa zq; do
  CMD2="python bench.py --scheme $sch --scenarios 8288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
  timeout 600 $CMD2 > $O/plain_$sch.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"solve_.*kernel" -s 6 -c 1 -f -o $O/prof_$sch $CMD2 > $O/ncu_$sch.log 2>&1
  echo "ncu $sch rc=$?" | tee -a $O/summary.txt
  python tools/ncu_summary.py $O/prof_$sch.ncu-rep $O/ncu_full_$sch.txt
  python tools/ncu_instmix.py $O/prof_$sch.ncu-rep 522144000 > $O/instmix_$sch.txt 2>&1
done
