#!/bin/bash
O=$PWD/gpurun_out/r2; mkdir -p $O
for lib in b200 scanonly minb8 w1e3; do
CRT1D_B200_LIB=$PWD/crt1d_b200/libcrt1d_$lib.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file $O/l_$lib.csv python bench.py --scheme 4s --scenarios 16576 --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --no-legs > /dev/null 2>&1
echo "== $lib"; grep -i "fixup\|solve_rows" $O/l_$lib.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -4
done
