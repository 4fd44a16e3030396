#!/bin/bash
# session 5, call E: source-level ncu capture of the closed-form zq flat kernel (rep comes back for reading here)
O=$PWD/gpurun_out/s5e; mkdir -p $O
CMD2="python bench.py --scheme zq --scenarios 8288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"solve_flat_kernel" -s 6 -c 1 -f -o $O/prof_zq_closed $CMD2 > $O/ncu_zq.log 2>&1
echo "ncu zq rc=$?"
ls -la $O
