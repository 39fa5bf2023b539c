"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel family and top launches."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot = collections.defaultdict(float)
cnt = collections.Counter()
rows = []
for i, row in enumerate(csv.DictReader(lines)):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(unit, 1e-6)
    name = row["Kernel Name"]
    short = re.sub(r"void |\(anonymous namespace\)::", "", name)
    m = re.match(r"([\w:]+)(<[^(]*>)?", short)
    short = (m.group(1) + (m.group(2) or "")) if m else short
    tot[short] += v
    cnt[short] += 1
    rows.append((i, short, v, row.get("Grid Size", ""), row.get("Block Size", "")))
T = sum(tot.values())
print(f"total {T:.3f} ms over {len(rows)} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:9.3f} ms {100 * v / T:5.1f}%  n={cnt[k]:3d}  {k[:110]}")
print("--- top single launches (index in launch order)")
for i, s, v, g, b in sorted(rows, key=lambda t: -t[2])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"#{i:3d} {v:8.3f} ms {100 * v / T:5.1f}%  grid {g:>14s} {s[:80]}")
