#!/usr/bin/env bash
# Builds the reference's binary decoder (gsw4869/CUDA_LDPC, bldpc_实习/) for the CPU from the
# sources where they lie under /root/reference, through the shim in oracle/shim/.  Outputs go
# ONLY into oracle/_ref/ (git-ignored; travels to the GPU box).  TEST INFRASTRUCTURE ONLY.
#
# On-the-fly edits of the source stream (nothing is copied into the repo):
#   K<<<g,b>>>(args);           -> SHIM_LAUNCH(K, g, b, args);       (serial kernel launch)
#   char file[100] = "PON_LDPC.txt" -> the config's H file (absolute path)
#   float Add_result;           -> float Add_result = 0;             (uninitialised)
#   variant "fixed" only: the Transform_H ternary's non-wrapping branch (B/Simulation.cu:380)
#   returns (Z-s)%Z+index3 instead of index3 — the one-line fix.
#
# usage: build_ref.sh NAME J L Z HFILE F MAXIT LEAST_TEST_FRAMES SNRTYPE [literal|fixed]
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
REF="${REF_ROOT:-/root/reference}/bldpc_实习"
[ -d "$REF" ] || { echo "reference tree not present ($REF) — keeping prebuilt oracle/_ref" >&2; exit 0; }
# REF_HFILE_DIR: directory the built library reads its H file from at run time (default: the
# reference tree; builds that must run on the GPU box point it at the repo's data copy).
HDIR="${REF_HFILE_DIR:-$REF}"
name=$1; J=$2; L=$3; Z=$4; hfile=$5; F=$6; maxit=$7; least=$8; snrtype=$9; variant=${10:-literal}
out="$here/_ref"; mkdir -p "$out"; tmp="$(mktemp -d)"; trap 'rm -rf "$tmp"' EXIT
CXX=/usr/bin/g++
defs="-DREF_J=$J -DREF_L=$L -DREF_Z=$Z -DREF_F=$F -DREF_MAXIT=$maxit -DREF_LEAST_TEST_FRAMES=$least -DREF_SNRTYPE=$snrtype"
flags="-O2 -fPIC -w -ffp-contract=off -include $here/shim/define_override.cuh -I $here/shim -I $REF $defs"
fix='s/__NOFIX__//'
[ "$variant" = fixed ] && fix='s/ - Z : index3;/ - Z : (Z - H[index1 * L + index0]) % Z + index3;/'
for src in Simulation LDPC_Decoder LDPC_Encoder; do
  tr -d '\r' < "$REF/$src.cu" \
   | sed -E 's/([A-Za-z_]+)<<<([^,]+),\s*([^>]+)>>>\((.*)\);/SHIM_LAUNCH(\1, \2, \3, \4);/' \
   | sed "s|char file\[100\] = \"PON_LDPC.txt\";|char file[200] = \"$HDIR/$hfile\";|" \
   | sed 's/float Add_result;/float Add_result = 0;/' \
   | sed "$fix" \
   | $CXX -x c++ $flags -c -o "$tmp/$src.o" -
done
$CXX $flags -c -o "$tmp/ref_entry.o" "$here/shim/ref_entry.cpp"
$CXX -shared -o "$out/libbldpc_ref_$name.so" "$tmp"/*.o -lm
echo "built $out/libbldpc_ref_$name.so"
