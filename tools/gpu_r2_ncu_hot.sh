#!/bin/bash
# ncu --set full + source page of the dominant kernel of each scheme in $SCHEMES: summary, instruction mix, hot lines
O=$PWD/gpurun_out/r2h; mkdir -p $O
T=$PWD/tools
for sch in ${SCHEMES:-4s zq n79}; do
  CMD="python bench.py --scheme $sch --scenarios 8288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-legs"
  timeout 600 $CMD > $O/plain_$sch.log 2>&1 || { echo "$sch plain run failed"; tail -5 $O/plain_$sch.log; continue; }
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"solve_(rows|flat|)_?kernel" -s 6 -c 1 -f -o /tmp/prof_$sch $CMD > $O/ncu_$sch.log 2>&1
  echo "ncu $sch rc=$?"
  python $T/ncu_summary.py /tmp/prof_$sch.ncu-rep $O/ncu_full_$sch.txt
  python $T/ncu_instmix.py /tmp/prof_$sch.ncu-rep 522144000 > $O/instmix_$sch.txt 2>&1
  python $T/ncu_hot.py /tmp/prof_$sch.ncu-rep $O/hot_$sch.txt 70
  head -3 $O/hot_$sch.txt
done
