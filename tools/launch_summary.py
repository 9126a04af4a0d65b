#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python tools/launch_summary.py profiles/r2_launches_train.csv [--last-step N_LAUNCHES]
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
data = []
for r in rows[hi + 1:]:
    if len(r) > mv:
        try:
            data.append((re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("<unnamed>::", ""), float(r[mv].replace(",", ""))))
        except ValueError:
            pass
if len(sys.argv) > 3 and sys.argv[2] == "--last-step":
    data = data[-int(sys.argv[3]):]
d = collections.defaultdict(lambda: [0, 0.0])
for name, v in data:
    d[name][0] += 1
    d[name][1] += v
tot = sum(v for _, v in data)
print(f"{len(data)} launches, {tot / 1e6:.3f} ms (serialised, cold-cache per-launch times)")
for k, (c, v) in sorted(d.items(), key=lambda x: -x[1][1])[:24]:
    print(f"{v / 1e6:9.3f} ms {c:5d} {100 * v / tot:5.1f}%  {k[:90]}")
