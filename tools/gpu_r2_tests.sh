#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee $O/summary.txt; tail -8 $O/pytest_gpu.log
