#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (time share of the step)."""
import collections, csv, re, sys

def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])[:72]
        v = float(r[vi].replace(",", "")); u = r[ui]
        v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0}.get(u, 1.0)
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# %s : %d launches, %.3f ms (ncu, cold-cache, serialised)" % (path, sum(v[0] for v in agg.values()), tot))
    print("%-74s %6s %10s %7s" % ("kernel", "n", "ms", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-74s %6d %10.3f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))

if __name__ == "__main__":
    main(sys.argv[1])
