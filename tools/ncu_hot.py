"""Top stall-sample lines of an `ncu --set full --import-source on` capture: tools/ncu_hot.py rep.ncu-rep out.txt [N]
Per SASS line: samples, share, times executed, the dominant stall reasons, the instruction."""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[h]
col = {n: i for i, n in enumerate(hdr)}
i_smp = col.get("# Samples", col.get("Warp Stall Sampling (All Samples)"))
stalls = [n for n in hdr if n.startswith("stall_") and "(Not Issued)" not in n]
recs, tot, by_reason = [], 0, {}
for r in rows[h + 1:]:
    try:
        e = int(r[col["Instructions Executed"]].replace(",", ""))
        s = int(r[i_smp].replace(",", "")) if r[i_smp] else 0
    except Exception:
        continue
    st = {}
    for n in stalls:
        try:
            v = int(r[col[n]].replace(",", "")) if r[col[n]] else 0
        except Exception:
            v = 0
        if v:
            st[n[6:]] = v
            by_reason[n[6:]] = by_reason.get(n[6:], 0) + v
    recs.append((s, e, r[col["Address"]][-5:], r[col["Source"]].strip()[:64], st))
    tot += s
with open(out, "w") as f:
    f.write("total samples %d\n" % tot)
    f.write("by reason: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, tot)) for k, v in sorted(by_reason.items(), key=lambda kv: -kv[1])[:8]) + "\n")
    for s, e, a, src, st in sorted(recs, key=lambda x: -x[0])[:top]:
        why = " ".join("%s=%d" % (k, v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        f.write("%7d %5.1f%% exe=%10d %s %-64s %s\n" % (s, 100.0 * s / max(1, tot), e, a, src, why))
