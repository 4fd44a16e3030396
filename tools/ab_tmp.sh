mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for sch in zq n79 zq_pa; do
timeout 300 python bench.py --scheme $sch --scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import json,sys
L=[l for l in sys.stdin if l.startswith('{')]
if not L: print('$sch FAILED'); sys.exit()
d=json.loads(L[-1]); r=d['roofline']
print('$sch value=%.3e frac=%.3f kernel_ms=%.3f' % (d['value'], r['frac'], r['kernel_ms']), d['clocks'].get('sm_mhz'))"
done
for sch in n79 zq; do
CMD2="python bench.py --scheme $sch --scenarios 8288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 $CMD2 > gpurun_out/plain_$sch.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 6 -c 1 -o gpurun_out/prof_$sch -f $CMD2 > gpurun_out/ncu_$sch.log 2>&1
echo "ncu $sch rc=$?"
done
