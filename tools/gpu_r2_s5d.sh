#!/bin/bash
# session 5, call D: zq closed form on equally spaced axes: gpu parity tests, A/B against the Thomas build (60 levels and n_z = 1000)
O=$PWD/gpurun_out/s5d; mkdir -p $O
line() { python - "$1" "$2" <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(sys.argv[2], "FAILED"); print(open(sys.argv[1]).read()[-600:]); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
print("%-28s value=%.4e frac=%.4f GB/s=%.0f kernel_ms=%.3f ms/step=%.2f chunk=%s sm_mhz=%s" % (sys.argv[2], d["value"], r["frac"], r["achieved"], r["kernel_ms"], d["ms_per_step"], d["config"].get("chunk"), c.get("sm_mhz")))
PY
}
timeout 900 python -m pytest tests -m gpu -q -x -k "zq or random or flat or deep or nonuniform or ragged or plugin or default" > $O/pytest_zq.log 2>&1; echo "pytest rc=$?" | tee $O/summary.txt; tail -3 $O/pytest_zq.log
S="--scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
D="--nz 1000 --scenarios 1184 --chunk 296 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
for rep in 1 2; do
for lib in default lib_thomas; do
  if [ "$lib" = default ]; then unset CRT1D_B200_LIB; else export CRT1D_B200_LIB=$PWD/_r1/$lib.so; fi
  timeout 300 python bench.py --scheme zq $S > $O/v.json 2> $O/v.err; line $O/v.json "zq $lib" | tee -a $O/summary.txt
  timeout 300 python bench.py --scheme zq $D > $O/v.json 2> $O/v.err; line $O/v.json "deep_zq $lib" | tee -a $O/summary.txt
done; done
unset CRT1D_B200_LIB
timeout 300 python bench.py --scheme zq_pa $S > $O/v.json 2> $O/v.err; line $O/v.json "zq_pa default(256thr)" | tee -a $O/summary.txt
timeout 300 python bench.py --scheme zq_pa $D > $O/v.json 2> $O/v.err; line $O/v.json "deep_zq_pa default" | tee -a $O/summary.txt
CMD2="python bench.py --scheme zq --scenarios 8288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"solve_.*kernel" -s 6 -c 1 -f -o /tmp/prof_zq $CMD2 > $O/ncu_zq.log 2>&1
echo "ncu zq rc=$?" | tee -a $O/summary.txt
python tools/ncu_summary.py /tmp/prof_zq.ncu-rep $O/ncu_full_zq.txt
python tools/ncu_instmix.py /tmp/prof_zq.ncu-rep 522144000 > $O/instmix_zq.txt 2>&1
