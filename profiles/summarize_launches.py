"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
idx = {h: i for i, h in enumerate(rows[0])}
agg, n = collections.OrderedDict(), collections.Counter()
for r in rows[1:]:
    try:
        v = float(r[idx["Metric Value"]])
    except ValueError:
        continue
    k = r[idx["Kernel Name"]][:90]
    agg[k] = agg.get(k, 0) + v
    n[k] += 1
tot = sum(agg.values())
print(f"# {sys.argv[1]}: {sum(n.values())} launches, {tot / 1e3:.1f} us total (cold-cache, serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda x: -x[1]):
    print(f"{v / 1e3:10.1f} us {100 * v / tot:5.1f}%  n={n[k]:5d}  avg={v / n[k] / 1e3:8.1f} us  {k}")
