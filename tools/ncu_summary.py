#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys


def summarise(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", "")) * scale.get(r[mu], 1.0)
    tot = sum(a[1] for a in agg.values())
    out = ["{:>11s} {:>6s} {:>6s}  {}".format("total ms", "share", "count", "kernel")]
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append("{:11.3f} {:5.1f}% {:6d}  {}   ({:.3f} ms/launch)".format(t, 100 * t / tot, c, n, t / c))
    out.append("{:11.3f} 100.0%".format(tot))
    return "\n".join(out)


if __name__ == "__main__":
    print(summarise(sys.argv[1]))
