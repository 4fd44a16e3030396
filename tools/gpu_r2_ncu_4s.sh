#!/bin/bash
O=$PWD/gpurun_out/r2; mkdir -p $O
T=$PWD/tools
cap() { name=$1; dir=$2; extra=$3
  (cd $dir && timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_rows -s 6 -c 1 -f -o /tmp/$name python bench.py --scheme 4s --scenarios 8288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e $extra > $O/ncu_$name.log 2>&1); echo "$name rc=$?"
  python $T/ncu_summary.py /tmp/$name.ncu-rep $O/ncu_$name.txt
  python $T/ncu_instmix.py /tmp/$name.ncu-rep 522144000 > $O/instmix_$name.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page source --csv > /tmp/$name.src.csv 2>/dev/null; python - /tmp/$name.src.csv $O/hot_$name.txt <<'PY'
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
def col(n):
    return hdr.index(n) if n in hdr else None
i_src=col("Source"); i_exe=col("Instructions Executed"); i_smp=col("# Samples") if col("# Samples") is not None else col("Warp Stall Sampling (All Samples)")
i_addr=col("Address")
out=[]
tot_s=0
for r in rows[2:]:
    try:
        e=int(r[i_exe].replace(",","")); s=int(r[i_smp].replace(",","")) if i_smp is not None and r[i_smp] else 0
    except Exception: continue
    out.append((s,e,r[i_addr] if i_addr is not None else "",r[i_src].strip()[:70])); tot_s+=s
with open(sys.argv[2],"w") as f:
    f.write("columns: %s\n"%hdr)
    f.write("total samples %d\n"%tot_s)
    for s,e,a,src in sorted(out,reverse=True)[:60]:
        f.write("%7d %5.1f%% exe=%10d %s %s\n"%(s,100.0*s/max(1,tot_s),e,a,src))
PY
}
cap 4s_now . --no-legs
cap 4s_r1 _r1 ""
