#!/bin/bash
# full GPU pass: gpu test tier, smoke, full bench (2s), per-scheme short benches, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
l=[x for x in open("gpurun_out/bench_full.log") if x.startswith("{")]
d=json.loads(l[-1]); r=d["roofline"]
print("2s FULL value=%.4e e2e=%.4e frac=%.3f GB/s=%.0f kernel_ms=%.3f cpu=%.3e (%d cores)" % (d["value"], d["e2e"]["value"], r["frac"], r["achieved"], r["kernel_ms"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]), d["clocks"])
PY
if [ -z "$SKIP_NCU" ]; then
CMD="python bench.py --scenarios 33152 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_launches.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_2s.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
CMD2="python bench.py --scenarios 16576 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 $CMD2 > gpurun_out/plain_full.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:solve_2s_rows -s 12 -c 2 -o gpurun_out/prof_2s_rows $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a gpurun_out/summary.txt
fi
