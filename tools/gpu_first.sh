#!/bin/bash
# first GPU pass: smoke, gpu tests, short + full bench (logs under gpurun_out/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --scenarios 65536 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_small.log 2>&1; echo "bench_small rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/bench_small.log
timeout 600 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench_full rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/bench_full.log
cat gpurun_out/smoke.log | tail -5
