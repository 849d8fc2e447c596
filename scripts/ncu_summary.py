#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (mean ns, share)."""
import collections
import csv
import sys


def main(path, per_step=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    grids = {}
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        name = name.replace("void ", "").replace("fcn::", "")
        key = name + " grid=" + row["Grid Size"].replace(" ", "")
        agg.setdefault(key, []).append(float(row["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print("%-78s %5s %12s %7s" % ("kernel", "n", "mean_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-78s %5d %12.2f %6.1f%%" % (k[:78], len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    print("total %.1f us over %d launches" % (tot / 1e3, sum(len(v) for v in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1])
