"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) as a markdown table."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
reader = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
for r in reader:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3
    name = r["Kernel Name"]
    agg[name][0] += 1
    agg[name][1] += us
total = sum(v[1] for v in agg.values())
print("| kernel | launches | total us | share | avg us |")
print("|---|---|---|---|---|")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name[:70]}` | {n} | {us:.1f} | {100 * us / total:.1f}% | {us / n:.1f} |")
