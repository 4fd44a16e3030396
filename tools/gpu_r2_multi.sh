#!/bin/bash
# N-GPU strong-scaling bench lines (driver-style launch); N from $NGPU (default 2); schemes from $SCHEMES
O=gpurun_out/r2; mkdir -p $O
N=${NGPU:-2}
for sch in ${SCHEMES:-2s 4s}; do
  for n in 1 $N; do
    if [ $n = 1 ]; then
      timeout 900 python bench.py --gpus 1 --scheme $sch --steps 3 --warmup 3 --no-cpu-baseline --host-sample 0 > $O/scale_${sch}_1.json 2> $O/scale_${sch}_1.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --scheme $sch --steps 3 --warmup 3 > $O/scale_${sch}_$n.json 2> $O/scale_${sch}_$n.err
    fi
    echo "$sch N=$n rc=$?"
    python - $O/scale_${sch}_$n.json <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print("NO LINE"); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]
print("  value=%.4e ms/step=%.2f frac=%.3f share=%.3f scaling=%s" % (d["value"], d["ms_per_step"], r["frac"], r["kernel_share_of_step"], d["scaling"]))
if "multi_gpu" in d: print("  multi:", json.dumps(d["multi_gpu"])[:700])
if "weak_scaling" in d: print("  weak:", d["weak_scaling"]["value"], d["weak_scaling"]["ms_per_step"])
PY
    tail -3 $O/scale_${sch}_$n.err
  done
done
