#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total and mean duration."""
import collections
import csv
import sys


def main(path, top=45):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi or not r[vi]:
            continue
        n = r[ki]
        n = n[:n.index("(")] if "(" in n and "<" not in n[:n.index("(")] else n.split("(const")[0].split("(Radix")[0].split("(Rows")[0]
        a = agg.setdefault(n.strip(), [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(t for _, t in agg.values())
    print("%d launches, %.1f us of kernel time" % (sum(c for c, _ in agg.values()), tot / 1e3))
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print("%-80s n=%4d tot=%9.1f us avg=%8.2f us %5.1f%%" % (k[:80], c, t / 1e3, t / c / 1e3, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
