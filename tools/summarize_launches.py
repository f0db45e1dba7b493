"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / mean, grouped by phase."""
import csv
import re
import sys
from collections import OrderedDict, defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
    rows.append((int(r["ID"]), name, ns))


def short(n):
    m = re.match(r"(?:void )?(?:vly::)?(\w+)(<[^>]*>)?", n)
    return (m.group(1) + (m.group(2) or "")) if m else n[:60]


agg = defaultdict(lambda: [0, 0.0])
for _, n, ns in rows:
    a = agg[short(n)]
    a[0] += 1
    a[1] += ns
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, total {tot / 1e3:.1f} us")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / 1e3:12.1f} us  {100 * t / tot:5.1f}%  n={c:5d}  mean={t / c / 1e3:9.2f} us  {k}")
