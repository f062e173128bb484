#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum CSV) per kernel: count, mean us, share."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg.setdefault(r[ki].split("(")[0][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | mean us | share of profiled time |\n|---|---|---|---|")
    for k, v in agg.items():
        print("| `%s` | %d | %.1f | %.1f %% |" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))


if __name__ == "__main__":
    main(sys.argv[1])
