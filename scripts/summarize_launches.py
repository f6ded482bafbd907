"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
    python scripts/summarize_launches.py gpurun_out/launches.csv [first_id last_id]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
tot = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    i = int(row["ID"])
    if not (lo <= i <= hi):
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void ", "", name)[:70]
    us = float(row["Metric Value"].replace(",", "")) / 1000.0
    c = tot.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += us
total = sum(v[1] for v in tot.values())
print(f"| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n} | {us:.1f} | {100 * us / total:.1f}% |")
print(f"| **sum** | {sum(v[0] for v in tot.values())} | {total:.1f} | 100% |")
