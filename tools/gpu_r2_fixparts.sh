#!/bin/bash
O=$PWD/gpurun_out/r2g; mkdir -p $O
for P in 1 2 4 8; do
CMD="python bench.py --scheme 4s --scenarios 16576 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-legs"
CRT1D_B200_FIXUP_PARTS=$P timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_4s_fp$P.csv $CMD > $O/ncu_4sfix.log 2>&1
echo "parts=$P"; grep -v "^==" $O/launches_4s_fp$P.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    if 'fixup' in r['Kernel Name']: print(r['Kernel Name'][:30], r['Grid Size'], r['Metric Value'], r['Metric Unit'])
" | tail -3
done
