#!/usr/bin/env bash
# C4 (LDPC_N576_K288_GF64, 64-QAM): FER with the all-zero codeword (= constellation point 0, a corner point: the
# reference's only option for this matrix) against random encoded codewords (nb_ldpc_encode, --encode with
# three seeds): how far does the corner-point bias move the curve?  Device channel, >= 200 frame errors per point.
S=./cuda_ldpc_b200/nb_ldpc_sim; D=cuda_ldpc_b200/data/nbldpc
C4="--matrix $D/LDPC_N576_K288_GF64_d1_exp.txt --exp --constellation $D/Constellation/GRAY_64QAM.txt"
run() { echo "== $1"; shift; $S "$@" --least-errors 200 --least-frames 20000 --max-frames 4000000 --batch 8192 2>&1 | grep -E "^ *[0-9-]" ; }
for algo in ems tmm; do
  run "C4 64-QAM $algo, all-zero codeword (reference CPU EMS: 8 dB 0.668, 9 dB 0.336, 10 dB 0.0791)" $C4 --algo $algo --snr 8 11 1
  for seed in 1 2 3; do run "C4 64-QAM $algo, random encoded codeword, seed $seed" $C4 --algo $algo --snr 8 11 1 --encode --seed $seed; done
done
rm -f /tmp/results_test.txt
$S $C4 --algo tmm --snr 10 10 1 --encode --max-frames 8192 --results /tmp/results_test.txt > /dev/null 2>&1; echo "== --results row (reference format, myNBLDPC/src/Simulation.cpp:205):"; cat /tmp/results_test.txt
