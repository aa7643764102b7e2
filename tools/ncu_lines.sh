#!/bin/bash
# per-CUDA-source-line executed instructions / stall samples of one NB decoder (run on the GPU box):
#   tools/ncu_lines.sh ems C4 2048  -> gpurun_out/lines_ems_C4.txt
rep=/tmp/lines_$1_$2
ncu --set full --clock-control none --import-source on -k regex:nb_decode -s 2 -c 1 -f -o $rep python tools/prof_nb.py $1 $2 $3 > /dev/null 2>&1
(cd /tmp && ncu -i $rep.ncu-rep --page source --csv --print-source cuda,sass > $rep.csv 2>/dev/null)
python tools/ncu_lines.py $rep.csv 70 > gpurun_out/lines_$1_$2.txt
rm -f $rep.ncu-rep $rep.csv
