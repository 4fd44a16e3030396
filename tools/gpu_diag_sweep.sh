#!/bin/bash
# reduced-diagnostic launch-configuration sweep + gpu tests + default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt; tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/diag_sweep.txt
run() { echo "== $1" >> gpurun_out/diag_sweep.txt; shift; timeout 300 env "$@" python tools/diag_only.py 2>&1 | grep scheme | cut -c1-75 >> gpurun_out/diag_sweep.txt; }
run "minb2 t128" X=1
run "minb2 t64" CRT1D_B200_DIAG_THREADS=64
run "minb2 t256" CRT1D_B200_DIAG_THREADS=256
run "minb3 t128" CRT1D_B200_LIB=$PWD/crt1d_b200/libcrt1d_b200_dg3.so
run "minb3 t256" CRT1D_B200_LIB=$PWD/crt1d_b200/libcrt1d_b200_dg3.so CRT1D_B200_DIAG_THREADS=256
run "minb4 t128" CRT1D_B200_LIB=$PWD/crt1d_b200/libcrt1d_b200_dg4.so
run "minb4 t256" CRT1D_B200_LIB=$PWD/crt1d_b200/libcrt1d_b200_dg4.so CRT1D_B200_DIAG_THREADS=256
cat gpurun_out/diag_sweep.txt
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
tail -c 1500 gpurun_out/bench_full.log
