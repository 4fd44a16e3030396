for sch in zq n79; do for lib in cur curv1 tri3 tri256 tri5 tri5v1; do
unset CRT1D_B200_LIB CRT1D_B200_FORCE_VEC1
case $lib in cur) ;; curv1) export CRT1D_B200_FORCE_VEC1=1;; tri5v1) export CRT1D_B200_FORCE_VEC1=1 CRT1D_B200_LIB=/root/repo/crt1d_b200/libcrt1d_b200_tri5.so;; *) export CRT1D_B200_LIB=/root/repo/crt1d_b200/libcrt1d_b200_$lib.so;; esac
timeout 300 python bench.py --scheme $sch --scenarios 66304 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import json,sys
L=[l for l in sys.stdin if l.startswith('{')]
if not L: print('$lib $sch FAILED'); sys.exit()
d=json.loads(L[-1]); r=d['roofline']
print('$lib $sch value=%.3e frac=%.3f kernel_ms=%.3f' % (d['value'], r['frac'], r['kernel_ms']), d['clocks'].get('sm_mhz'))"
done; done
