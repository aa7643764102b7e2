#!/bin/bash
# one `ncu --set full` capture of the fp16 message mode's kernel (run on the GPU box AFTER the plain command exited 0):
#   tools/ncu_layered_f16.sh <tag> [configs]  -> gpurun_out/<tag>_ncu_layered_f16_<C>.txt
tag=${1:-r02}
cfgs=${2:-C2}
declare -A F=( [C1]=32768 [C2]=1184 [C3]=8192 )
declare -A EG=( [C1]=$((32768/2*7680*10)) [C2]=$((1184/2*147200*10)) [C3]=$((8192/2*70400*10)) )
for c in $cfgs; do
  python tools/prof_one.py $c ${F[$c]} 10 3 fp16 > gpurun_out/${tag}_plain_f16_$c.log 2>&1 || { echo "plain run of $c failed"; exit 1; }
  rep=/tmp/${tag}_f16_$c
  ncu --set full --clock-control none --import-source on -k regex:ldpc_layered_f16 -s 2 -c 1 -f -o $rep \
      python tools/prof_one.py $c ${F[$c]} 10 3 fp16 > gpurun_out/${tag}_ncu_f16_$c.log 2>&1
  echo "$c ncu rc=$?"
  out=gpurun_out/${tag}_ncu_layered_f16_$c.txt
  { echo "# ncu --set full --clock-control none: ldpc_layered_f16_kernel, $c, ${F[$c]} frames, 10 iterations fixed, msg_max 31, x0.875 (tools/ncu_layered_f16.sh)";
    python tools/ncu_summary.py $rep.ncu-rep;
    (cd /tmp && ncu -i $rep.ncu-rep --page source --csv --print-source sass > $rep.csv 2>/dev/null);
    echo "-- executed warp instructions by opcode: percent, per edge and 2 codewords (top 24)"; python tools/ncu_ophist.py $rep.csv ${EG[$c]} 2>/dev/null | head -25; } > $out 2>&1
  rm -f $rep.ncu-rep $rep.csv
done
