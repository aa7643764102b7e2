#!/usr/bin/env bash
# Builds the reference's non-binary simulator (gsw4869/CUDA_LDPC, myNBLDPC/src/*.cu, *.cpp — ALL sources, main.cu
# included, compiled UNCHANGED from where they lie under /root/reference) FOR THE GPU (sm_100) with the configuration
# macros of baseline/ref_gpu/nb_define_override_gpu.h in place of include/define.h.  Output: baseline/_ref/nbldpc_gpu_<name>
# (git-ignored; travels to the GPU box).  It reads its matrix / GF table / constellation relative to the working
# directory (tools/run_ref_gpu.sh prepares one) and appends to results.txt there, like the original.
# usage: build_ref_nb_gpu.sh NAME MATRIX CONSTFILE NQAM GFQ MAXDC MAXDV NM NC MAXIT METHOD SNR LEAST_ERRORS LEAST_FRAMES
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
REF="${REF_ROOT:-/root/reference}/myNBLDPC"
[ -d "$REF" ] || { echo "reference tree not present ($REF) — keeping prebuilt baseline/_ref" >&2; exit 0; }
name=$1; matrix=$2; cfile=$3; nqam=$4; gfq=$5; maxdc=$6; maxdv=$7; nm=$8; nc=$9; maxit=${10}; method=${11}; snr=${12}; le=${13}; lf=${14}
out="$here/_ref"; mkdir -p "$out"; tmp="$(mktemp -d)"; trap 'rm -rf "$tmp"' EXIT
NVCC=/usr/local/cuda/bin/nvcc
defs="-DNBREF_MATRIX=\"$matrix\" -DNBREF_CONST=\"$cfile\" -DNBREF_NQAM=$nqam -DNBREF_GFQ=$gfq -DNBREF_MAXDC=$maxdc -DNBREF_MAXDV=$maxdv -DNBREF_NM=$nm -DNBREF_NC=$nc -DNBREF_MAXIT=$maxit -DNBREF_METHOD=$method -DNBREF_SNR=$snr -DNBREF_LEAST_ERRORS=$le -DNBREF_LEAST_FRAMES=$lf"
flags="-O3 -w -gencode arch=compute_100,code=sm_100 -ccbin /usr/bin/g++ -Xcompiler -pthread -include $here/ref_gpu/nb_define_override_gpu.h -I $REF/include"
objs=""
for src in main.cu Decode_GPU.cu LDPC_Decoder.cpp GF.cpp LDPC_Encoder.cpp Simulation.cpp struct.cpp; do
  # shellcheck disable=SC2086
  $NVCC $flags $defs -x cu -dc -o "$tmp/$src.o" "$REF/src/$src"
  objs="$objs $tmp/$src.o"
done
# shellcheck disable=SC2086
$NVCC -gencode arch=compute_100,code=sm_100 -ccbin /usr/bin/g++ -Xcompiler -pthread -o "$out/nbldpc_gpu_$name" $objs
echo "built $out/nbldpc_gpu_$name"
