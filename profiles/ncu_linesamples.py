"""Stall samples per source line (top N) of an .ncu-rep captured with --import-source on.
usage: python profiles/ncu_linesamples.py report.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
cur = None
samp = collections.Counter()
inst = collections.Counter()
src = {}
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split('/')[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10:
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    extra = len(r) - len(hdr)
    try:
        samp[(cur, ln)] += int(r[hdr.index("# Samples") + extra])
        inst[(cur, ln)] += int(r[hdr.index("Instructions Executed") + extra])
    except ValueError:
        pass
    src[(cur, ln)] = ",".join(r[1:2 + extra])[:90]
tot = max(1, sum(samp.values()))
print("samples", tot, "warp instructions", sum(inst.values()))
for k, v in samp.most_common(top):
    print("%-18s %4d smp %6d %5.1f%%  inst %9d  %s" % (k[0], k[1], v, 100.0 * v / tot, inst[k], src[k]))
