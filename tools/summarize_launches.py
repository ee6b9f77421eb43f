"""Summarise an ncu launch list (gpu__time_duration.sum) per kernel name: count, total, share."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if unit in ("ns", "nsecond"):
        v /= 1000.0
    elif unit in ("ms", "msecond"):
        v *= 1000.0
    tot[name][0] += 1
    tot[name][1] += v
total = sum(v[1] for v in tot.values())
print(f"total kernel time {total/1000:.3f} ms over {sum(v[0] for v in tot.values())} launches")
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t/1000:9.3f} ms {100*t/total:5.1f}%  n={n:5d}  avg={t/n:8.1f} us  {name[:110]}")
