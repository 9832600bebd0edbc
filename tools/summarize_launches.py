#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [> profiles/xyz.md]"""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0][-60:]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v * 1e6 if u in ("s", "second") else v
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1] / 1e3:.3f} | {a[1] / a[0]:.1f} | {100 * a[1] / tot:.1f}% |")
    print(f"\ntotal {tot / 1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches (ncu serialised, cold cache: compare shares)")


if __name__ == "__main__":
    main(sys.argv[1])
