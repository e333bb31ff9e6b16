"""Aggregate the last step of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv) per kernel.
Usage: python scripts/summarize_launches.py launches.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = rows[hi + 1:]
names = [r[idx['Kernel Name']].split('(')[0].replace('slotvps::', '').replace('void ', '') for r in data]
vals = [float(r[idx['Metric Value']].replace(',', '')) for r in data]
mi = [i for i, n in enumerate(names) if 'mask_tc' in n]
L = mi[-1] - mi[-2]                                   # launches per step (mask_tc runs once per step)
# the step ends with its last fuse_relabel launch (the rows timed after the step -- tracker, id-map consumers, the deformable-conv
# subnet -- follow it in the list): take the L launches that end there
ends = [i for i, n in enumerate(names) if 'fuse_relabel' in n]
end = (ends[-1] + 1) if ends else len(names)
agg = collections.OrderedDict(); tot = 0.0
for i in range(end - L, end):
    a = agg.setdefault(names[i], [0, 0.0]); a[0] += 1; a[1] += vals[i] / 1000.0; tot += vals[i] / 1000.0
print(f"launches per step {L}, serialised kernel time {tot:.1f} us (cold cache, one launch at a time)")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n[:44]:44s} x{c:3d} {t:8.1f} us  {t / c:7.1f} us each  {100 * t / tot:5.1f} %")
