"""SASS opcode histogram per kernel of the built library (cuobjdump -sass): how many DMMA / 256-bit stores / cp.async
(LDGSTS) / TMA bulk copies (UBLKCP) / cluster instructions each kernel really contains.
Usage: python tools/sass_histogram.py [lib] > profiles/r2_sass_histogram.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "bpp_b200/libbppgpu.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
KEY = ("DMMA", "DFMA", "DMUL", "DADD", "STG", "LDG", "LDGSTS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDS", "STS",
       "SHFL", "BAR", "UCGABAR", "MEMBAR", "ERRBAR", "CCTL", "MUFU", "ATOMS", "VOTE")
fn = None
hist = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        hist[fn][op] += 1
for fn in sorted(hist):
    h = hist[fn]
    total = sum(h.values())
    name = demangle(fn)
    name = re.sub(r"bppgpu::", "", name)
    if len(name) > 110:
        name = name[:107] + "..."
    groups = collections.Counter()
    for op, c in h.items():
        base = op.split(".")[0]
        if base in KEY:
            groups[base] += c
        if op.startswith("STG") and "256" in op:
            groups["STG.256"] += c
        if op.startswith("LDG") and "256" in op:
            groups["LDG.256"] += c
        if op.startswith("DMMA"):
            groups[op] += 0
    print("%-112s %6d instr  " % (name, total) + "  ".join("%s=%d" % (k, v) for k, v in sorted(groups.items()) if v))
