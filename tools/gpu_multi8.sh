#!/bin/bash
NG=${NG:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus $NG --steps 2 --warmup 3 > gpurun_out/bench_n$NG.log 2>&1; echo "bench n$NG rc=$?"
python - <<PY
import json
f = "gpurun_out/bench_n$NG.log"
l=[x for x in open(f) if x.startswith("{")]
if not l:
    print("NO JSON"); print(open(f).read()[-2500:])
else:
    d=json.loads(l[-1])
    print("n_gpus", d["n_gpus"], "value %.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], "ms/step %.1f" % d["ms_per_step"], d.get("clocks"))
PY
