#!/bin/bash
# The reference's own NON-BINARY simulator with its GPU decoders (baseline/_ref/nbldpc_gpu_*, built by baseline/build_all.sh)
# on the GPU box, and this repo's nb_ldpc_sim on the same code / modulation / decoder / Eb/N0 / transmitted word beside it.
#   tools/run_ref_nb_gpu.sh [out_file]
out=${1:-gpurun_out/r02_reference_nb_gpu.txt}
root=$(pwd); D=$root/cuda_ldpc_b200/data/nbldpc; R=$root/baseline/_ref; S=$root/cuda_ldpc_b200/nb_ldpc_sim
w=$(mktemp -d); mkdir -p $w/GF
ln -s $D/Constellation $w/Constellation
for f in $D/*.txt; do ln -s $f $w/; done
python - "$w" <<'PY'
import json, os, sys
sys.path.insert(0, os.getcwd())
from cuda_ldpc_b200.gf import write_table_file
w = sys.argv[1]
for q in (16, 64, 256):
    write_table_file(q, os.path.join(w, "GF", f"Arith.Table.GF.{q}.txt"))
cw = json.load(open("tests/golden/nb_ref.json"))["CodeWord_sym_test"]
open(os.path.join(w, "cw96.txt"), "w").write(" ".join(map(str, cw)))
open(os.path.join(w, "cw72.txt"), "w").write(" ".join(map(str, cw[:72])))
PY
: > "$out"
run_ref() { echo "== reference GPU decoder: $1" >> "$out"; (cd $w && rm -f results.txt && timeout 600 $R/nbldpc_gpu_$2 > log.txt 2>&1; echo "rc=$?" >> "$root/$out"; grep -E "^ *-?[0-9]+\.[0-9] " log.txt >> "$root/$out"); }
run_us() { echo "== this repo (nb_ldpc_sim): $1" >> "$out"; shift; (cd $w && timeout 300 $S "$@" 2>&1 | grep -E "^ *-?[0-9]+\.[0-9] " >> "$root/$out"); }
echo "# columns: Eb/N0 frames errFrames FER SER avgIter sec/frame [info Mbit/s]" >> "$out"
run_ref "BDS GF64 BPSK EMS(2,2), 2.5 dB, CodeWord_sym_test (the committed configuration)" BDS_ems
run_us  "same" --matrix BDS.576.288.GF.64.txt --constellation Constellation/BPSK.txt --algo ems --snr 2.5 2.5 1 --codeword cw96.txt --least-errors 200 --least-frames 20000 --batch 8192
run_ref "BDS GF64 BPSK TMM, 2.5 dB" BDS_tmm
run_us  "same" --matrix BDS.576.288.GF.64.txt --constellation Constellation/BPSK.txt --algo tmm --snr 2.5 2.5 1 --codeword cw96.txt --least-errors 200 --least-frames 20000 --batch 8192
run_ref "C4 GF64 64-QAM EMS(2,2), 10 dB, raw exponents as coefficients and a non-codeword (as committed): 20 iterations on every frame" C4_ems
run_us  "C4 64-QAM EMS(2,2), 10 dB, the same transmitted word (not a codeword: 20 iterations on every frame)" --matrix LDPC_N576_K288_GF64_d1_exp.txt --exp --constellation Constellation/GRAY_64QAM.txt --algo ems --snr 10 10 1 --codeword cw96.txt --max-frames 16384 --batch 8192
run_ref "C5 GF256 BPSK TMM, 4.5 dB, non-codeword: 20 iterations on every frame" C5_tmm
run_us  "C5 TMM, 4.5 dB, same transmitted word (20 iterations on every frame)" --matrix LDPC_N576_K480_GF256_exp.txt --exp --constellation Constellation/BPSK.txt --algo tmm --snr 4.5 4.5 1 --codeword cw72.txt --max-frames 8192 --batch 4096
cat "$out"; rm -rf $w
