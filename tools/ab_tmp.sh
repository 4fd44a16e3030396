timeout 600 python -m pytest tests -m gpu -q -x -k "zq or n79" 2>&1 | tail -3
for sch in zq n79 zq_pa; do for lib in cur ck5 ck6 ck10; do
if [ $lib = cur ]; then unset CRT1D_B200_LIB; else export CRT1D_B200_LIB=/root/repo/crt1d_b200/libcrt1d_b200_$lib.so; fi
if [ $sch = zq_pa ] && [ $lib != cur ]; then continue; fi
timeout 300 python bench.py --scheme $sch --scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import json,sys
L=[l for l in sys.stdin if l.startswith('{')]
if not L: print('$lib $sch FAILED'); sys.exit()
d=json.loads(L[-1]); r=d['roofline']
print('$lib $sch value=%.3e frac=%.3f kernel_ms=%.3f' % (d['value'], r['frac'], r['kernel_ms']), d['clocks'].get('sm_mhz'))"
done; done
