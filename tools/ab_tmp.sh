timeout 600 python -m pytest tests -m gpu -q -x -k "zq_pa or deep or variant or indexing" 2>&1 | tail -3
for sch in zq_pa; do for lib in cur pa2 pa4v1 pa4 pa4ck6; do
if [ $lib = cur ]; then unset CRT1D_B200_LIB; else export CRT1D_B200_LIB=/root/repo/crt1d_b200/libcrt1d_b200_$lib.so; fi
timeout 300 python bench.py --scheme $sch --scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import json,sys
L=[l for l in sys.stdin if l.startswith('{')]
if not L: print('$lib $sch FAILED'); sys.exit()
d=json.loads(L[-1]); r=d['roofline']
print('$lib $sch value=%.3e frac=%.3f kernel_ms=%.3f' % (d['value'], r['frac'], r['kernel_ms']), d['clocks'].get('sm_mhz'))"
done; done
