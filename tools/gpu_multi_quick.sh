#!/bin/bash
# N-GPU bench lines (weak scaling) for 2s (full sweep) and 4s / zq (132 608 scenarios per GPU), plus N=1 on the same box
NG=${NG:-2}
mkdir -p gpurun_out; : > gpurun_out/multi.txt
run() {  # scheme scenarios ngpu
  if [ $3 = 1 ]; then
    timeout 900 python bench.py --gpus 1 --scheme $1 --scenarios $2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/v.log 2>&1
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $3 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $3 --scheme $1 --scenarios $2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/v.log 2>&1
  fi
  python - "$1" "$3" <<'PY' | tee -a gpurun_out/multi.txt
import json, sys
l=[x for x in open("gpurun_out/v.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1])
    print("%-4s n_gpus=%d value=%.4e e2e=%.4e ms/step=%.1f scaling=%s clocks=%s" % (sys.argv[1], d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["scaling"], d["clocks"].get("sm_mhz")))
else:
    print(sys.argv[1], sys.argv[2], "FAILED"); print(open("gpurun_out/v.log").read()[-1200:])
PY
}
for n in 1 $NG; do run 2s 1000000 $n; run 4s 132608 $n; run zq 132608 $n; done
