#!/usr/bin/env python
"""Developer tool: top SASS lines by warp-stall samples from `ncu -i rep --page source --csv` output.
usage: ncu_hot.py source.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = rows[1]
ci = {k: i for i, k in enumerate(h)}
data = rows[2:]
tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[ci[k]] or 0) for r in data) for k in stalls}
print("total samples", tot)
print({k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
idx = sorted(range(len(data)), key=lambda i: -int(data[i][ci["# Samples"]] or 0))[:n]
for i in sorted(idx):
    r = data[i]
    top = sorted(((int(r[ci[k]] or 0), k) for k in stalls), reverse=True)[:3]
    print(f"{i:5d} {r[ci['Address']][-6:]} {int(r[ci['# Samples']]):6d} {r[ci['Source']][:70]:70s} " +
          " ".join(f"{k[6:]}={v}" for v, k in top if v))
