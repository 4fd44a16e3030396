#!/bin/bash
# round-2 first GPU pass: gpu test tier, smoke, full bench (2s), reference arm
O=gpurun_out/r2; mkdir -p $O
nproc > $O/host.txt; free -g >> $O/host.txt; nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv >> $O/host.txt
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee $O/summary.txt; tail -5 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt; tail -1 $O/smoke.log
( time timeout 900 python bench.py > $O/bench_1.json 2> $O/bench_1.err ); echo "bench rc=$?" | tee -a $O/summary.txt; tail -3 $O/bench_1.err
( time timeout 900 python bench.py --impl reference > $O/bench_ref_1.json 2> $O/bench_ref_1.err ); echo "ref rc=$?" | tee -a $O/summary.txt
cat $O/bench_ref_1.json
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/r2/bench_1.json") if x.startswith("{")][-1]); r=d["roofline"]
print("2s value=%.4e e2e=%.4e frac=%.3f kernel_ms=%.3f" % (d["value"], d["e2e"]["value"], r["frac"], r["kernel_ms"]))
for k in ("cpu_baseline","e2e_host_profiles","reduced_diagnostic","nonuniform_lai","clocks"): print(k, json.dumps(d.get(k))[:900])
PY
