#!/bin/bash
# round-2 FINAL GPU pass on the committed binary (session 5): gpu test tier, smoke, full bench (2s) + reference arm, ncu launch list,
# ncu --set full of the 2s and 4s row-sweep kernels, one bench line per scheme, deep canopy.  Outputs: gpurun_out/r2g/
O=$PWD/gpurun_out/r2g; mkdir -p $O
T=$PWD/tools
line() { python - "$1" "$2" <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(sys.argv[2], "FAILED"); print(open(sys.argv[1]).read()[-600:]); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
print("%-10s value=%.4e frac=%.4f GB/s=%.0f kernel_ms=%.3f ms/step=%.2f sm_mhz=%s reasons=%s kernel=%s" % (sys.argv[2], d["value"], r["frac"], r["achieved"], r["kernel_ms"], d["ms_per_step"], c.get("sm_mhz"), c.get("reasons"), r.get("kernel")))
PY
}
if [ -z "$SKIP_TESTS" ]; then
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee $O/summary.txt; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt; tail -1 $O/smoke.log
fi
if [ -z "$SKIP_BENCH" ]; then
timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?" | tee -a $O/summary.txt
timeout 900 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "bench rc=$?" | tee -a $O/summary.txt
line $O/bench_full.json 2s_full | tee -a $O/summary.txt
fi
if [ -z "$SKIP_SCHEMES" ]; then
: > $O/all_schemes.txt
for sch in 2s bl 4s bf g77 zq n79 zq_pa; do
  timeout 300 python bench.py --scheme $sch --scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > $O/v_$sch.json 2> $O/v.err
  line $O/v_$sch.json $sch | tee -a $O/all_schemes.txt
done
: > $O/deep.txt
for sch in 4s 2s zq n79 zq_pa bf; do
  timeout 600 python bench.py --scheme $sch --nz 1000 --scenarios 1184 --chunk -296 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > $O/d_$sch.json 2> $O/v.err
  line $O/d_$sch.json "deep_$sch" | tee -a $O/deep.txt
done
fi
if [ -z "$SKIP_NCU" ]; then
CMD="python bench.py --scenarios 33152 --steps 2 --warmup 3 --no-cpu-baseline --no-legs"
timeout 600 $CMD > $O/plain_launches.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_2s.csv $CMD > $O/ncu_launches.log 2>&1
echo "ncu launches rc=$?" | tee -a $O/summary.txt
for sch in ${NCU_SCHEMES:-2s 4s zq n79 zq_pa}; do
  CMD2="python bench.py --scheme $sch --scenarios 8288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
  timeout 600 $CMD2 > $O/plain_$sch.log 2>&1 && \
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"solve_.*kernel" -s 6 -c 1 -f -o /tmp/prof_$sch $CMD2 > $O/ncu_$sch.log 2>&1
  echo "ncu $sch rc=$?" | tee -a $O/summary.txt
  python $T/ncu_summary.py /tmp/prof_$sch.ncu-rep $O/ncu_full_$sch.txt
  python $T/ncu_instmix.py /tmp/prof_$sch.ncu-rep 522144000 > $O/instmix_$sch.txt 2>&1
done
fi
