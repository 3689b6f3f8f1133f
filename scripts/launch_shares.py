"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` log by kernel name: launches, total and average time.
   python scripts/launch_shares.py gpurun_out/launches_b4.csv [name-regex-to-keep]"""
import collections
import csv
import re
import sys

lines = open(sys.argv[1]).read().splitlines(True)
keep = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines[start:]):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1e-3)
    k = re.sub(r"\(.*", "", row["Kernel Name"]).strip()
    if keep is not None and not keep.search(k):
        continue
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v[1] / tot * 100:7.2f}% {v[0]:6d} {v[1] / v[0]:10.2f} us  {k}")
