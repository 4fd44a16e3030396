#!/bin/bash
# GPU pass 2: full gpu test tier, per-scheme throughput, ncu launch list + full capture of the 2s kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
for sch in 4s bf g77 bl zq n79; do
  timeout 600 python bench.py --scheme $sch --scenarios 131072 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$sch.log 2>&1
  echo "bench $sch rc=$?" | tee -a gpurun_out/summary.txt
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench_$sch.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); r=d["roofline"]
    print("$sch", "value=%.3e"%d["value"], "frac=%.3f"%r["frac"], "GB/s=%.0f"%r["achieved"], "kernel_ms=%.3f"%r["kernel_ms"], d["clocks"])
PY
done
CMD="python bench.py --scenarios 32768 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_launches.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_2s.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
CMD2="python bench.py --scenarios 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 $CMD2 > gpurun_out/plain_full.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 12 -c 2 -o gpurun_out/prof_2s $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out
