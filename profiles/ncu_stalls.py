"""Top SASS instructions of one kernel by stall samples, with the dominant stall reasons.
Usage: python profiles/ncu_stalls.py report.ncu-rep kernel_regex [top] [reason]"""
import csv
import subprocess
import sys


def main():
    rep, kn = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    only = sys.argv[4] if len(sys.argv) > 4 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                          "--kernel-name", "regex:" + kn], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    recs = []
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        try:
            smp = int(d["# Samples"])
        except ValueError:
            continue
        st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit()}
        recs.append((smp, d["Address"], d["Source"], st, int(d["Instructions Executed"] or 0)))
    tot = sum(r[0] for r in recs)
    agg = {}
    for r in recs:
        for k, v in r[3].items():
            agg[k] = agg.get(k, 0) + v
    print("samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    key = (lambda r: -r[3].get(only, 0)) if only else (lambda r: -r[0])
    for smp, addr, src, st, ie in sorted(recs, key=key)[:top]:
        dom = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print("%6d %5.1f%% %s  %-60s %s" % (smp, 100.0 * smp / max(1, tot), addr[-5:], src[:60],
                                          " ".join("%s=%d" % kv for kv in dom if kv[1])))


if __name__ == "__main__":
    main()
