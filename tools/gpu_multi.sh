#!/bin/bash
# multi-GPU validation: gpu tests on rank-0 GPU, then bench at N=1 and N=$NG via torchrun
NG=${NG:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary_multi.txt; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; echo "bench n1 rc=$?" | tee -a gpurun_out/summary_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 2 --warmup 3 > gpurun_out/bench_n$NG.log 2>&1; echo "bench n$NG rc=$?" | tee -a gpurun_out/summary_multi.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $NG --steps 1 --warmup 1 --cpu-sample 256 > gpurun_out/bench_ref_n$NG.log 2>&1; echo "bench ref n$NG rc=$?" | tee -a gpurun_out/summary_multi.txt
python - <<PY
import json
for f in ("gpurun_out/bench_n1.log", "gpurun_out/bench_n$NG.log", "gpurun_out/bench_ref_n$NG.log"):
    l=[x for x in open(f) if x.startswith("{")]
    if not l:
        print(f, "NO JSON"); print(open(f).read()[-1500:]); continue
    d=json.loads(l[-1])
    print(f, "n_gpus", d["n_gpus"], "value %.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], "ms/step %.1f" % d["ms_per_step"], d.get("clocks"))
PY
