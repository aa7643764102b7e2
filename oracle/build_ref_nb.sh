#!/usr/bin/env bash
# Builds the reference's non-binary CPU decoder (gsw4869/CUDA_LDPC, myNBLDPC/src/*.cpp) from the
# sources where they lie under /root/reference into oracle/_ref/libnbldpc_ref_<name>.so (git-ignored).
# The sources are compiled UNCHANGED (g++ on the files directly); only the configuration macros
# come from oracle/shim/nb_define_override.h and the CUDA runtime header is the CPU shim.
# TEST INFRASTRUCTURE ONLY.
# usage: build_ref_nb.sh NAME MATRIX CONSTFILE NQAM GFQ MAXDC MAXDV NM NC MAXIT
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
REF="${REF_ROOT:-/root/reference}/myNBLDPC"
[ -d "$REF" ] || { echo "reference tree not present ($REF) — keeping prebuilt oracle/_ref" >&2; exit 0; }
name=$1; matrix=$2; cfile=$3; nqam=$4; gfq=$5; maxdc=$6; maxdv=$7; nm=$8; nc=$9; maxit=${10}
out="$here/_ref"; mkdir -p "$out"; tmp="$(mktemp -d)"; trap 'rm -rf "$tmp"' EXIT
CXX=/usr/bin/g++
defs="-DNBREF_MATRIX=\"$matrix\" -DNBREF_CONST=\"$cfile\" -DNBREF_NQAM=$nqam -DNBREF_GFQ=$gfq -DNBREF_MAXDC=$maxdc -DNBREF_MAXDV=$maxdv -DNBREF_NM=$nm -DNBREF_NC=$nc -DNBREF_MAXIT=$maxit -DNBREF_METHOD=0"
flags="-O2 -fPIC -w -ffp-contract=off -pthread -include $here/shim/nb_define_override.h -I $here/shim -I $REF/include"
for src in LDPC_Decoder GF LDPC_Encoder Simulation struct; do
  # shellcheck disable=SC2086
  $CXX $flags $defs -c -o "$tmp/$src.o" "$REF/src/$src.cpp"
done
# shellcheck disable=SC2086
$CXX $flags $defs -c -o "$tmp/nb_ref_entry.o" "$here/shim/nb_ref_entry.cpp"
$CXX -shared -pthread -o "$out/libnbldpc_ref_$name.so" "$tmp"/*.o -lm
echo "built $out/libnbldpc_ref_$name.so"
