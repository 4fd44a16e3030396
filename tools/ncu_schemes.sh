mkdir -p gpurun_out
for sch in $NCU_SCHEMES; do
  CMD="python bench.py --scheme $sch --scenarios 8288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
  timeout 600 $CMD > gpurun_out/plain_$sch.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"solve_.*kernel" -s 6 -c 1 -f -o gpurun_out/prof_$sch $CMD > gpurun_out/ncu_$sch.log 2>&1
  echo "ncu $sch rc=$?"
done
