"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/ncu_summary.py launches.csv [> summary.md]"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, ns))
agg = OrderedDict()
for name, ns in rows:
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
total = sum(v[1] for v in agg.values())
print(f"| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {cnt} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% |")
print(f"| **total** | {len(rows)} | {total / 1e6:.3f} | 100% |")
