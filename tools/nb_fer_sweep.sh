#!/usr/bin/env bash
# NB FER points of BASELINE.md §2 (the reference's own CPU decoder, seeds 173, 50 error frames) re-run on the device
# channel (Philox) through nb_ldpc_sim with many more frames: same code / modulation / decoder / Eb/N0.
S=./cuda_ldpc_b200/nb_ldpc_sim; D=cuda_ldpc_b200/data/nbldpc
run() { echo "== $1"; shift; $S "$@" --least-errors 200 --least-frames 20000 --max-frames 2000000 --batch 8192 2>&1 | grep -E "^ *[0-9-]" ; }
run "BDS GF64 BPSK EMS(2,2)   reference: 2.0 dB 0.797, 2.5 dB 0.363, 3.0 dB 0.058" --matrix $D/BDS.576.288.GF.64.txt --constellation $D/Constellation/BPSK.txt --algo ems --snr 2.0 3.0 0.5
run "BDS GF64 BPSK TMM        reference: 2.0 dB 0.563, 3.0 dB 0.0447" --matrix $D/BDS.576.288.GF.64.txt --constellation $D/Constellation/BPSK.txt --algo tmm --snr 2.0 3.0 1.0
run "BDS GF64 BPSK layered    reference: 2.0 dB 0.0196" --matrix $D/BDS.576.288.GF.64.txt --constellation $D/Constellation/BPSK.txt --algo ltmm --snr 2.0 2.0 1.0
run "C4 GF64 64QAM EMS(2,2)   reference: 8 dB 0.668, 9 dB 0.336, 10 dB 0.0791" --matrix $D/LDPC_N576_K288_GF64_d1_exp.txt --exp --constellation $D/Constellation/GRAY_64QAM.txt --algo ems --snr 8 10 1
run "C5 GF256 BPSK TMM        reference: 4 dB 0.265, 5 dB 0.00559" --matrix $D/LDPC_N576_K480_GF256_exp.txt --exp --constellation $D/Constellation/BPSK.txt --algo tmm --snr 4 5 1
run "C5 GF256 BPSK FFT-BP     (no reference decoder; must not be worse than TMM)" --matrix $D/LDPC_N576_K480_GF256_exp.txt --exp --constellation $D/Constellation/BPSK.txt --algo fftbp --snr 4 5 1
run "C5 GF256 BPSK EMS(2,2)   reference: 4 dB 0.940, 5 dB 0.405" --matrix $D/LDPC_N576_K480_GF256_exp.txt --exp --constellation $D/Constellation/BPSK.txt --algo ems --snr 4 5 1
