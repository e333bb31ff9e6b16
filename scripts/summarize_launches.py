#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (device time, share)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
body = [r for r in rows[hdr + 1:] if len(r) > vi]
# one full step = the launches after the second-to-last fuse_relabel (last kernel of a step) up to the last one
ends = [i for i, r in enumerate(body) if "fuse_relabel" in r[ki]]
if len(ends) >= 2 and "--all" not in sys.argv:
    body = body[ends[-2] + 1:ends[-1] + 1]
    print(f"# one step: launches {ends[-2] + 1}..{ends[-1]} of the capture")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in body:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("slotvps::", "")[:60]
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] == "us" else v
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':52s} {'launches':>8s} {'ms':>9s} {'share':>7s} {'us/launch':>10s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:52s} {v[0]:8d} {v[1]:9.3f} {100 * v[1] / tot:6.1f}% {1e3 * v[1] / v[0]:10.1f}")
print(f"{'total':52s} {sum(v[0] for v in agg.values()):8d} {tot:9.3f}")
