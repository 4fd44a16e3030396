#!/bin/bash
# strong scaling (BASELINE.json configs[3]: one 10^6-scenario sweep block-partitioned over NG ranks) for 2s and 4s
NG=${NG:-8}
O=$PWD/gpurun_out/r2s; mkdir -p $O
for sch in 2s 4s; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus $NG --scheme $sch --steps 2 --warmup 3 --no-cpu-baseline --no-legs > $O/scale_${sch}_$NG.json 2> $O/scale_${sch}_$NG.err; echo "bench $sch n$NG rc=$?"
python - $O/scale_${sch}_$NG.json <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l:
    print("NO JSON"); print(open(sys.argv[1].replace(".json",".err")).read()[-2500:])
else:
    d=json.loads(l[-1]); m=d.get("multi_gpu",{})
    print(d["config"]["scheme"], "n_gpus", d["n_gpus"], d["scaling"], "value %.4e" % d["value"], "ms/step %.2f" % d["ms_per_step"], "kernels/step per rank", ["%.1f"%x for x in m.get("per_rank_kernels_ms_per_step",[])], "collective_exposed_ms_total", m.get("collective_exposed_ms_total"))
PY
done
