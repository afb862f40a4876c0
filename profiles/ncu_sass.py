"""SASS listing of one kernel from an .ncu-rep with stall samples and executed counts per instruction.
Usage: python profiles/ncu_sass.py report.ncu-rep [min_samples]   (prints hot instructions in order)"""
import csv
import subprocess
import sys

rep = sys.argv[1]
mins = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = None
tot = sum(int(r[ismp]) for r in rows[2:] if len(r) > iex and r[ismp].isdigit())
for n, r in enumerate(rows[2:]):
    if len(r) <= iex or not r[ismp].isdigit():
        continue
    a = int(r[ia], 16)
    base = a if base is None else base
    if int(r[ismp]) >= mins:
        print("%5d %05x %6d %5.2f%% %9s  %s" % (n, a - base, int(r[ismp]), 100.0 * int(r[ismp]) / max(1, tot), r[iex], r[isrc].strip()[:90]))
