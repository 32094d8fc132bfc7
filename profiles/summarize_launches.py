#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python profiles/summarize_launches.py gpurun_out/launches.csv profiles/r02_ncu_launches_summary.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[col["Metric Value"]].replace(",", "").replace(".", "").isdigit():
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).strip()
    v = float(r[col["Metric Value"]].replace(",", ""))
    unit = r[col["Metric Unit"]]
    us = v / 1000 if unit.startswith("ns") else v * 1000 if unit.startswith("ms") else v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
out = [("kernel", "launches", "total_us", "mean_us", "share_of_captured_time")]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append((k, n, round(t, 1), round(t / n, 2), round(t / tot, 4)))
csv.writer(open(sys.argv[2], "w")).writerows(out)
for o in out[:12]:
    print(o)
