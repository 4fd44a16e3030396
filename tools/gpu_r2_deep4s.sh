#!/bin/bash
O=$PWD/gpurun_out/r2g; mkdir -p $O
CMD="python bench.py --scheme 4s --nz 1000 --scenarios 592 --chunk 296 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-legs"
timeout 300 $CMD > $O/d4s_plain.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_deep4s.csv $CMD > $O/ncu_d4s.log 2>&1
grep -v "^==" $O/launches_deep4s.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['Kernel Name'][:70], r['Grid Size'], r['Block Size'], r['Metric Value'], r['Metric Unit'])
" | tail -24
