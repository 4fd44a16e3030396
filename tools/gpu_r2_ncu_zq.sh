#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
CMD="python bench.py --scheme zq --order matrix --scenarios 16576 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
cap() { name=$1; k=$2; shift 2
  env "$@" timeout 900 ncu --set full --clock-control none -k regex:$k -s 5 -c 1 -o /tmp/$name $CMD > $O/ncu_$name.log 2>&1; echo "$name rc=$?"
  python tools/ncu_summary.py /tmp/$name.ncu-rep $O/ncu_$name.txt
}
cap zq_multi_v1 solve_multi CRT1D_B200_FORCE_VEC1=1
cap zq_multi_v2 solve_multi A=1
cap zq_single_v2 solve_kernel CRT1D_B200_MULTI=0
cap zq_single_v1 solve_kernel CRT1D_B200_MULTI=0 CRT1D_B200_FORCE_VEC1=1
