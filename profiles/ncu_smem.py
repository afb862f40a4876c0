"""Shared-memory wavefronts per source line of one kernel from an .ncu-rep.
Usage: python profiles/ncu_smem.py report.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, cur = None, None
W, I, N, src = collections.Counter(), collections.Counter(), collections.Counter(), {}
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1]
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
    e = len(r) - len(hdr)
    try:
        W[(cur, ln)] += int(r[hdr.index("L1 Wavefronts Shared") + e])
        I[(cur, ln)] += int(r[hdr.index("L1 Wavefronts Shared Ideal") + e])
        N[(cur, ln)] += int(r[hdr.index("Instructions Executed") + e])
    except ValueError:
        pass
    src[(cur, ln)] = ",".join(r[1:2 + e])[:90]
t = max(1, sum(W.values()))
print("total shared wavefronts %d, instructions %d" % (t, sum(N.values())))
for k, v in W.most_common(top):
    print("%-16s %4d %10d %5.1f%% ideal %10d inst %9d  %s" % (k[0].split("/")[-1], k[1], v, 100.0 * v / t, I[k], N[k], src[k]))
