#!/bin/bash
# gpu tests + reduced-diagnostic launch-configuration sweep + FP64 peak microbenchmark + ncu of the 2s DIAG kernel + default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt; tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/diag_sweep.txt
run() { echo "== $1" >> gpurun_out/diag_sweep.txt; shift; timeout 300 env "$@" python tools/diag_only.py 2>&1 | grep scheme | cut -c1-75 >> gpurun_out/diag_sweep.txt; }
run "default" X=1
cp gpurun_out/diag_only.json gpurun_out/diag_only_default.json
run "t64" CRT1D_B200_DIAG_THREADS=64
run "t96" CRT1D_B200_DIAG_THREADS=96
run "t32" CRT1D_B200_DIAG_THREADS=32
cat gpurun_out/diag_sweep.txt
./tools/micro/fp64peak | tee gpurun_out/fp64peak.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_2s_rows -s 20 -c 1 -f -o gpurun_out/prof_2s_diag python tools/diag_only.py 2s > gpurun_out/ncu_diag.log 2>&1; echo "ncu diag rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
tail -c 1800 gpurun_out/bench_full.log
