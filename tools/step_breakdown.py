#!/usr/bin/env python
"""Per-kernel breakdown of the last full step in an ncu gpu__time_duration launch list.
usage: python tools/step_breakdown.py launches.csv <first-kernel-substring>"""
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in csv.DictReader(lines)
        if r.get("Metric Name") == "gpu__time_duration.sum"]
idx = [i for i, (n, _) in enumerate(rows) if sys.argv[2] in n]
a, b = idx[-2], idx[-1]
tot = sum(t for _, t in rows[a:b])
for n, t in rows[a:b]:
    nm = re.sub(r"\(.*", "", n).replace("<unnamed>::", "").replace("void ", "")[:90]
    print(f"{t/1000:8.2f} us  {100*t/tot:5.1f}%  {nm}")
print(f"total {tot/1000:.1f} us over {b-a} launches")
