#!/bin/bash
# session 5, call A: new tests (global level tables, preferred batch), deep-canopy A/B (whole-wave chunks, n79 tables in
# global memory), zq_pa at 80 registers (3 x 256-thread register bound -> 5 CTAs of 128 threads per SM)
O=$PWD/gpurun_out/s5a; mkdir -p $O
line() { python - "$1" "$2" <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(sys.argv[2], "FAILED"); print(open(sys.argv[1]).read()[-600:]); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
print("%-28s value=%.4e frac=%.4f GB/s=%.0f kernel_ms=%.3f ms/step=%.2f chunk=%s sm_mhz=%s kernel=%s" % (sys.argv[2], d["value"], r["frac"], r["achieved"], r["kernel_ms"], d["ms_per_step"], d["config"].get("chunk"), c.get("sm_mhz"), r.get("kernel","")[:40]))
PY
}
timeout 900 python -m pytest tests -m gpu -q -x -k "preferred or global_level or deep_canopy or flat_column" > $O/pytest_new.log 2>&1; echo "pytest rc=$?" | tee $O/summary.txt; tail -3 $O/pytest_new.log
D="--nz 1000 --scenarios 1184 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
for rep in 1 2; do
  timeout 300 python bench.py --scheme zq $D --chunk 296 > $O/v.json 2> $O/v.err; line $O/v.json "deep_zq chunk296" | tee -a $O/summary.txt
  timeout 300 python bench.py --scheme zq $D --chunk -296 > $O/v.json 2> $O/v.err; line $O/v.json "deep_zq whole-waves" | tee -a $O/summary.txt
  CRT1D_B200_NO_GTAB=1 timeout 300 python bench.py --scheme n79 $D --chunk 296 > $O/v.json 2> $O/v.err; line $O/v.json "deep_n79 smem-tables c296" | tee -a $O/summary.txt
  timeout 300 python bench.py --scheme n79 $D --chunk 296 > $O/v.json 2> $O/v.err; line $O/v.json "deep_n79 gtab c296" | tee -a $O/summary.txt
  timeout 300 python bench.py --scheme n79 $D --chunk -296 > $O/v.json 2> $O/v.err; line $O/v.json "deep_n79 gtab whole-waves" | tee -a $O/summary.txt
done
S="--scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
for rep in 1 2; do
  timeout 300 python bench.py --scheme zq_pa $S > $O/v.json 2> $O/v.err; line $O/v.json "zq_pa default" | tee -a $O/summary.txt
  CRT1D_B200_LIB=$PWD/_r1/lib_zqpa3.so timeout 300 python bench.py --scheme zq_pa $S > $O/v.json 2> $O/v.err; line $O/v.json "zq_pa 80regs" | tee -a $O/summary.txt
done
timeout 300 python bench.py --scheme zq $S > $O/v.json 2> $O/v.err; line $O/v.json "zq c4144" | tee -a $O/summary.txt
timeout 300 python bench.py --scheme zq $S --chunk -4144 > $O/v.json 2> $O/v.err; line $O/v.json "zq whole-waves" | tee -a $O/summary.txt
timeout 300 python bench.py --scheme n79 $S > $O/v.json 2> $O/v.err; line $O/v.json "n79 c4144" | tee -a $O/summary.txt
timeout 300 python bench.py --scheme n79 $S --chunk -4144 > $O/v.json 2> $O/v.err; line $O/v.json "n79 whole-waves" | tee -a $O/summary.txt
