#!/usr/bin/env python
"""Developer tool: executed-instruction histogram by opcode from `ncu -i rep --page source --csv --print-source sass`.
usage: ncu_ophist.py source.csv [edge_groups]   (edge_groups = frames/4 * E * iters: prints warp-inst*32 per edge-group)"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
iS, iI = hdr.index("Source"), hdr.index("Instructions Executed")
ops = collections.Counter(); tot = 0
for r in data:
    try: n = int(r[iI])
    except ValueError: continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
    if m: ops[m.group(2)] += n; tot += n
eg = float(sys.argv[2]) if len(sys.argv) > 2 else None
print("total warp-inst", tot, ("= %.2f per edge and 4 codewords" % (tot * 32 / eg)) if eg else "")
for k, v in ops.most_common(40):
    print(f"{v / tot * 100:6.2f}%  {(v * 32 / eg) if eg else 0:6.2f}  {k}")
