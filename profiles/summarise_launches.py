#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean us, share."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        agg[name].append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:72]:72s} n={len(v):4d} avg={sum(v)/len(v):9.1f} us  share={100*sum(v)/tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
