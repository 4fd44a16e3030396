"""Instruction mix per layer.band unit from an .ncu-rep captured with --import-source on (read here, no GPU needed).

    python tools/ncu_instmix.py gpurun_out/prof_zq.ncu-rep <units per launch>
"""
import collections
import csv
import io
import subprocess
import sys


def main(rep, units):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    i_src, i_exe = hdr.index("Source"), hdr.index("Instructions Executed")
    by, tot = collections.Counter(), 0
    for r in rows[2:]:
        try:
            n = int(r[i_exe].replace(",", ""))
        except (ValueError, IndexError):
            continue
        t = r[i_src].strip().split()
        op = t[1] if t[0].startswith("@") else t[0]
        op = op.rstrip(";")
        key = op if op.startswith(("IMAD", "MOV")) else op.split(".")[0]
        by[key] += n
        tot += n
    wunits = units / 32.0
    print(f"{rows[0][1][:90]}\n  {tot / wunits:.1f} thread-instructions per unit")
    for op, n in by.most_common(18):
        print(f"  {op:16s} {n / wunits:6.1f} per unit  {100.0 * n / tot:5.1f} %")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]))
