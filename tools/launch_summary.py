"""Per-kernel launch statistics of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_summary.py gpurun_out/launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ci = {h: i for i, h in enumerate(rows[hi])}
d = collections.OrderedDict()
for r in rows[hi + 2:]:
    if len(r) > ci["Metric Value"]:
        d.setdefault(r[ci["Kernel Name"]].split("(")[0], []).append(
            float(r[ci["Metric Value"]].replace(",", "")) / 1000.0)
tot = sum(sum(v) for v in d.values())
print(f"{'kernel':28s} {'n':>4s} {'avg us':>9s} {'min us':>9s} {'sum us':>10s} {'share':>6s}")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:28s} {len(v):4d} {sum(v) / len(v):9.2f} {min(v):9.2f} {sum(v):10.1f} {sum(v) / tot:6.1%}")
