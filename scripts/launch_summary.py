"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`) into the
per-kernel table kept under profiles/:  python scripts/launch_summary.py gpurun_out/X.csv profiles/<name>.csv"""
import csv
import re
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ix = {k: j for j, k in enumerate(rows[h])}
agg = OrderedDict()
for r in rows[h + 1:]:
    if len(r) != len(rows[h]) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[r[ix["Metric Unit"]]]
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).strip()
    n, t = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, t + v)
total = sum(t for _, t in agg.values())
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "total_us (ncu gpu__time_duration, cold-cache serialised)", "share"])
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([name, n, round(t, 1), round(t / total, 3)])
    w.writerow(["TOTAL", sum(n for n, _ in agg.values()), round(total, 1), 1.0])
print(open(sys.argv[2]).read())
