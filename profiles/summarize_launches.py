#!/usr/bin/env python
"""Per-kernel time table of ONE bench step from an `ncu --metrics gpu__time_duration.sum` launch list
(cold-cache, serialised launches: compare shares, not absolutes).
usage: summarize_launches.py launches.csv [step_index]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
names = [r[4].split("(")[0].replace("<unnamed>::", "").replace("void ", "") for r in rows]
starts = [i for i, n in enumerate(names) if n.startswith("filter_count_kernel")]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1
a, b = starts[k], (starts[k + 1] if k + 1 < len(starts) else len(rows))
agg, tot = collections.OrderedDict(), 0.0
for n, r in zip(names[a:b], rows[a:b]):
    t = float(r[14])
    tot += t
    e = agg.setdefault(n, [0, 0.0])
    e[0] += 1
    e[1] += t
print(f"step {k}: {b - a} launches, sum of kernel durations {tot / 1e3:.1f} us")
print(f"{'us':>10} {'share':>6} {'n':>3}  kernel")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t / 1e3:10.1f} {100 * t / tot:5.1f}% {c:3d}  {n}")
