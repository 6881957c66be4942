#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_summary.py launches.csv"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) <= mi:
            continue
        try:
            v = float(r[mi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = r[ki].split("(")[0][-60:]
        if "<" in r[ki]:
            name = r[ki][:r[ki].index("(", r[ki].index(">"))][-70:] if "(" in r[ki][r[ki].index(">"):] else r[ki][-70:]
        agg[name][0] += 1
        agg[name][1] += v * scale
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':72s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:72s} {v[0]:8d} {v[1]:10.3f} {100 * v[1] / tot:6.1f}%")
    print(f"{'total':72s} {sum(v[0] for v in agg.values()):8d} {tot:10.3f}")


if __name__ == "__main__":
    main()
