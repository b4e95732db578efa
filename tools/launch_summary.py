#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total device
time and share per kernel.  usage: tools/launch_summary.py gpurun_out/launches.csv"""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[row["Metric Unit"]]
        a = agg.setdefault(row["Kernel Name"], [0, 0.0, row["Block Size"], row["Grid Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"| launches | total us | share | us/launch | block | grid (last) | kernel |\n|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        name = k.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        print(f"| {a[0]} | {a[1]:.1f} | {a[1] / tot * 100:.2f}% | {a[1] / a[0]:.1f} | {a[2]} | {a[3]} | `{name[:80]}` |")


if __name__ == "__main__":
    main(sys.argv[1])
