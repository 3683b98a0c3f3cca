#!/usr/bin/env python
"""Executed-instruction histogram along the SASS of one kernel of an `ncu --set full --import-source on` report.

usage: python profiles/sass_histogram.py <report.ncu-rep> <kernel-name-regex> [bin=80]

Prints, per bin of consecutive SASS instructions, the executed warp instructions (share of the launch), the average active threads
and the dominant opcodes -- enough to attribute the instruction budget to the phases of a kernel (probe / scan / merge ...)."""
import csv, io, subprocess, sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    width = int(sys.argv[3]) if len(sys.argv) > 3 else 80
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-count", "1"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    isrc, ie, iavg = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
    body = [r for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
    tot = sum(int(r[ie]) for r in body)
    print(rows[0][1][:100], "| SASS instructions", len(body), "| executed", round(tot / 1e6, 2), "M warp-instructions")
    for k in range(0, len(body), width):
        seg = body[k:k + width]
        e = sum(int(r[ie]) for r in seg)
        if e < 0.01 * tot:
            continue
        ops = {}
        for r in seg:
            t = r[isrc].split()
            op = t[1] if t[0].startswith("@") else t[0]
            ops[op] = ops.get(op, 0) + int(r[ie])
        top = sorted(ops.items(), key=lambda x: -x[1])[:5]
        thr = sum(float(r[iavg]) * int(r[ie]) for r in seg) / max(e, 1)
        print(f"{k:5d}  {e / 1e6:6.2f} M ({100 * e / tot:4.1f} %)  thr {thr:4.1f}  " + " ".join(f"{a}:{b / 1e6:.2f}" for a, b in top))


if __name__ == "__main__":
    main()
