#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
per kernel name: launches, total time, share, DRAM bytes.   python profiles/launch_list_summarise.py launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ix = {k: i for i, k in enumerate(h)}
per = collections.OrderedDict()
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
byid = collections.defaultdict(dict)
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    name = re.sub(r"void pcd::pcd_kernel<pcd::K|, pcd::\w+>\(.*|\(int\)|void pcd::\w+::|\(.*", "", r[ix["Kernel Name"]])
    v = float(r[ix["Metric Value"]].replace(",", "")) * scale.get(r[ix["Metric Unit"]], 1.0)
    byid[r[ix["ID"]]]["name"] = name
    byid[r[ix["ID"]]][r[ix["Metric Name"]]] = v
for d in byid.values():
    p = per.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    p[0] += 1
    p[1] += d.get("gpu__time_duration.sum", 0.0)
    p[2] += d.get("dram__bytes_read.sum", 0.0)
    p[3] += d.get("dram__bytes_write.sum", 0.0)
T = sum(v[1] for v in per.values())
print(f"# {len(byid)} launches, total {T / 1e3:.2f} ms (cold-cache, serialised: compare SHARES), DRAM read {sum(v[2] for v in per.values()):.0f} MB "
      f"write {sum(v[3] for v in per.values()):.0f} MB")
for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:34s} n={v[0]:3d} total_us={v[1]:8.1f} share={v[1] / T:.3f} dram_rd_MB={v[2]:8.1f} dram_wr_MB={v[3]:8.1f}")
