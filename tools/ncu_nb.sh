#!/bin/bash
# one `ncu --set full` capture of nb_decode_kernel per (algorithm, config) — run on the GPU box after the plain command
# exited 0.  The reports are summarised ON THE BOX (tools/ncu_summary.py + tools/ncu_ophist.py) and deleted: four
# full reports exceed what gpurun copies back.
tag=${1:-r02}
for spec in "ems C4 2048" "ltmm C4 2048" "tmm C5 1024" "fftbp C5 1024"; do
  set -- $spec
  python tools/prof_nb.py $1 $2 $3 > gpurun_out/${tag}_nbplain_$1_$2.log 2>&1 || { echo "plain run $spec failed"; continue; }
  rep=/tmp/${tag}_nb_$1_$2
  ncu --set full --clock-control none --import-source on -k regex:nb_decode -s 2 -c 1 -f -o $rep \
      python tools/prof_nb.py $1 $2 $3 > gpurun_out/${tag}_nbncu_$1_$2.log 2>&1
  echo "$spec ncu rc=$?"
  out=gpurun_out/${tag}_ncu_nb_$1_$2.txt
  { echo "# ncu --set full --clock-control none: nb_decode_kernel, $1 on $2, $3 frames, max 20 iterations (tools/ncu_nb.sh)";
    python tools/ncu_summary.py $rep.ncu-rep;
    (cd /tmp && ncu -i $rep.ncu-rep --page source --csv --print-source sass > $rep.csv 2>/dev/null);
    echo "-- executed warp instructions by opcode (top 22)"; python tools/ncu_ophist.py $rep.csv | head -23; } > $out 2>&1
  rm -f $rep.ncu-rep $rep.csv
done
