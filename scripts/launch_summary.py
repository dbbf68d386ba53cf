#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total time, share."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, out, note=""):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ik]).split("::")[-1]
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6}.get(r[iu], 1.0)
        c = agg.setdefault(name, [0, 0.0])
        c[0] += 1
        c[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        if note:
            f.write("# " + note + "\n")
        f.write("kernel,launches,total_us,share\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{n},{t:.1f},{t / tot:.4f}\n")


if __name__ == "__main__":
    main(*sys.argv[1:4])
