#!/bin/bash
# session 5, call O: full gpu tier + smoke on the binary with the 4-way 4s band split; 4s lines at 60 / 1000 levels
O=$PWD/gpurun_out/s5o; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee $O/summary.txt; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt; tail -1 $O/smoke.log
line() { python - "$1" "$2" <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(sys.argv[2], "FAILED"); print(open(sys.argv[1]).read()[-600:]); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]; c=d["clocks"]
print("%-10s value=%.4e frac=%.4f GB/s=%.0f kernel_ms=%.3f ms/step=%.2f sm_mhz=%s reasons=%s kernel=%s" % (sys.argv[2], d["value"], r["frac"], r["achieved"], r["kernel_ms"], d["ms_per_step"], c.get("sm_mhz"), c.get("reasons"), r.get("kernel")))
PY
}
timeout 300 python bench.py --scheme 4s --scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > $O/v_4s.json 2> $O/v.err; line $O/v_4s.json 4s | tee -a $O/summary.txt
timeout 300 python bench.py --scheme 4s --nz 1000 --scenarios 1184 --chunk -296 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > $O/d_4s.json 2> $O/v.err; line $O/d_4s.json deep_4s | tee -a $O/summary.txt
timeout 600 python bench.py --scheme 4s --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > $O/full_4s.json 2> $O/v.err; line $O/full_4s.json 4s_1e6 | tee -a $O/summary.txt
