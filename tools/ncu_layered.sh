#!/bin/bash
# one `ncu --set full` capture of the layered int8 kernel per config (run on the GPU box AFTER the plain command exited 0):
#   tools/ncu_layered.sh <tag>   -> gpurun_out/<tag>_{c1,c2,c3}.ncu-rep
tag=${1:-r02}
declare -A F=( [C1]=65536 [C2]=2368 [C3]=16384 )
for c in C2 C1 C3; do
  python tools/prof_one.py $c ${F[$c]} 10 3 > gpurun_out/${tag}_plain_$c.log 2>&1 || { echo "plain run of $c failed"; exit 1; }
  lc=$(echo $c | tr A-Z a-z)
  ncu --set full --clock-control none --import-source on -k regex:ldpc_layered_i8 -s 2 -c 1 -f -o gpurun_out/${tag}_$lc \
      python tools/prof_one.py $c ${F[$c]} 10 3 > gpurun_out/${tag}_ncu_$c.log 2>&1
  echo "$c ncu rc=$?"
done
