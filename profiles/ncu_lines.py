"""Per-source-line instruction counts and stall samples of one kernel from an .ncu-rep
(ncu --set full --import-source on).  Usage: python profiles/ncu_lines.py report.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    kn = ["--kernel-name", sys.argv[3]] if len(sys.argv) > 3 else []
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + kn,
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, cur = None, None
    inst, samp, src = collections.Counter(), collections.Counter(), {}
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
        extra = len(r) - len(hdr)          # commas inside the source text split it into extra fields
        try:
            inst[(cur, ln)] += int(r[hdr.index("Instructions Executed") + extra])
            samp[(cur, ln)] += int(r[hdr.index("# Samples") + extra])
        except ValueError:
            pass
        src[(cur, ln)] = ",".join(r[1:2 + extra])[:100]
    ti, ts = sum(inst.values()), max(1, sum(samp.values()))
    print("total warp instructions %d, samples %d" % (ti, ts))
    for k, v in inst.most_common(top):
        print("%-18s %4d %10d %5.1f%% smp %4.1f%%  %s" % (k[0].split("/")[-1], k[1], v, 100.0 * v / ti,
                                                        100.0 * samp[k] / ts, src[k]))


if __name__ == "__main__":
    main()
