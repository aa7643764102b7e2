#!/bin/bash
# one `ncu --set full` capture of the layered int8 kernel per config (run on the GPU box AFTER the plain command exited 0);
# summaries are written on the box (the three full reports together exceed what gpurun copies back):
#   tools/ncu_layered.sh <tag>   -> gpurun_out/<tag>_ncu_layered_i8_{C1,C2,C3}.txt, gpurun_out/<tag>_ncu_layered_i8.json
tag=${1:-r02}
declare -A F=( [C1]=65536 [C2]=2368 [C3]=16384 )
declare -A EG=( [C1]=$((65536/4*7680*10)) [C2]=$((2368/4*147200*10)) [C3]=$((16384/4*70400*10)) )
for c in C2 C1 C3; do
  python tools/prof_one.py $c ${F[$c]} 10 3 > gpurun_out/${tag}_plain_$c.log 2>&1 || { echo "plain run of $c failed"; exit 1; }
  rep=/tmp/${tag}_$c
  ncu --set full --clock-control none --import-source on -k regex:ldpc_layered_i8 -s 2 -c 1 -f -o $rep \
      python tools/prof_one.py $c ${F[$c]} 10 3 > gpurun_out/${tag}_ncu_$c.log 2>&1
  echo "$c ncu rc=$?"
  out=gpurun_out/${tag}_ncu_layered_i8_$c.txt
  { echo "# ncu --set full --clock-control none: ldpc_layered_i8_kernel, $c, ${F[$c]} frames, 10 iterations fixed, msg_max 31, x0.875 (tools/ncu_layered.sh)";
    python tools/ncu_summary.py $rep.ncu-rep --json gpurun_out/${tag}_ncu_layered_i8.json $c ${F[$c]};
    (cd /tmp && ncu -i $rep.ncu-rep --page source --csv --print-source sass > $rep.csv 2>/dev/null);
    echo "-- executed warp instructions by opcode: percent, per edge and 4 codewords (top 24)"; python tools/ncu_ophist.py $rep.csv ${EG[$c]} 2>/dev/null | head -25; } > $out 2>&1
  [ "$c" = C2 ] && cp $rep.ncu-rep gpurun_out/${tag}_c2.ncu-rep
  rm -f $rep.ncu-rep $rep.csv
done
