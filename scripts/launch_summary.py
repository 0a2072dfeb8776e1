#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: mean / last duration per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ki][:70], []).append(float(r[vi].replace(",", "")))
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:70s} n={len(v):3d} mean={sum(v) / len(v) / 1e3:9.1f} us  last={v[-1] / 1e3:9.1f}")
