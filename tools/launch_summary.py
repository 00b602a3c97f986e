"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_summary.py gpurun_out/<launches>.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
ix = {h: i for i, h in enumerate(rows[0])}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ix["Metric Unit"]], 1e-6)
    k = r[ix["Kernel Name"]][:72]
    agg[k][0] += 1
    agg[k][1] += v * scale
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:74s} launches {v[0]:4d}  {v[1]:10.3f} ms  {v[1] / tot * 100:5.1f}%")
