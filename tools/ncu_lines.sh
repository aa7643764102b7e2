#!/bin/bash
# per-CUDA-source-line executed instructions / stall samples of one NB decoder (run on the GPU box):
#   tools/ncu_lines.sh ems C4 2048  -> gpurun_out/lines_ems_C4.csv
rep=/tmp/lines_$1_$2
ncu --set full --clock-control none --import-source on -k regex:nb_decode -s 2 -c 1 -f -o $rep python tools/prof_nb.py $1 $2 $3 > /dev/null 2>&1
(cd /tmp && ncu -i $rep.ncu-rep --page source --csv --print-source cuda > $rep.csv 2>/dev/null)
python - "$rep.csv" > gpurun_out/lines_$1_$2.txt <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
# find header row containing "Instructions Executed"
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
h = rows[hi]; data = rows[hi + 1:]
ci = {k: i for i, k in enumerate(h)}
def num(r, k):
    try: return int(r[ci[k]])
    except Exception: return 0
tot = sum(num(r, "Instructions Executed") for r in data)
samp = sum(num(r, "# Samples") for r in data)
print("total inst", tot, "samples", samp)
keyline = "#" if "#" in ci else h[0]
out = sorted(data, key=lambda r: -num(r, "# Samples"))[:60]
for r in out:
    print(f"{r[0]:>6s} inst {num(r,'Instructions Executed')/max(tot,1)*100:5.2f}%  samples {num(r,'# Samples')/max(samp,1)*100:5.2f}%  {r[ci['Source']][:150] if 'Source' in ci else ''}")
PY
rm -f $rep.ncu-rep $rep.csv
