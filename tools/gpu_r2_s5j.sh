#!/bin/bash
# session 5, call J: driver-style N = 2 launch of both arms on the final binary (strong scaling default)
O=$PWD/gpurun_out/s5j; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/ref_n2.json 2> $O/ref_n2.err; echo "ref N=2 rc=$?" | tee $O/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench N=2 rc=$?" | tee -a $O/summary.txt
python - $O/bench_n2.json <<'PY' | tee -a $O/summary.txt
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print("NO LINE"); sys.exit()
d=json.loads(l[-1]); r=d["roofline"]
print("value=%.4e ms/step=%.2f frac=%.3f n_gpus=%d scaling=%s e2e=%.4e" % (d["value"], d["ms_per_step"], r["frac"], d["n_gpus"], d["scaling"], d["e2e"]["value"]))
print("multi:", json.dumps(d.get("multi_gpu"))[:600])
PY
tail -c 300 $O/ref_n2.json
