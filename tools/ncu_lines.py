#!/usr/bin/env python
"""Per-CUDA-source-line executed instructions and stall samples from `ncu -i rep --page source --csv --print-source cuda,sass`.
usage: ncu_lines.py source_cuda_sass.csv [N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
inst = collections.Counter(); samp = collections.Counter(); src = {}
cur_file = ""; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No" and "Instructions Executed" in r:
        hdr = {k: i for i, k in enumerate(r)}; iS = r.index("Source"); continue
    if hdr is None or r[0] == "Function Name": continue
    try:
        key = (cur_file, int(r[0]))
    except ValueError:
        continue
    if r[iS].strip(): src[key] = r[iS].strip()
    try:
        inst[key] += int(r[hdr["Instructions Executed"]] or 0); samp[key] += int(r[hdr["# Samples"]] or 0)
    except (ValueError, IndexError):
        pass
ti, ts = sum(inst.values()), sum(samp.values())
print("total inst", ti, "samples", ts)
for key, v in samp.most_common(n):
    print(f"{key[0]}:{key[1]:<5d} inst {inst[key] / max(ti, 1) * 100:5.2f}%  samples {v / max(ts, 1) * 100:5.2f}%  {src.get(key, '')[:130]}")
