#!/usr/bin/env bash
# Builds the reference's binary decoder (gsw4869/CUDA_LDPC, bldpc_实习/) FOR THE GPU (sm_100) from the sources where
# they lie under /root/reference, with the per-config constants of SURVEY Appendix A force-included in place of the
# committed (inconsistent) define.cuh, and baseline/ref_gpu/ref_gpu_harness.cu in place of main.cu.
# Outputs go ONLY into baseline/_ref/ (git-ignored; travels to the GPU box).  The reference's own build system
# (CMake hard-wired to /usr/local/cuda-11.0) is not run.
#
# On-the-fly edit of the source stream (nothing is copied into the repo): the H file name hard-coded in Get_H
# (B/Simulation.cu:296).  Variant "fixed" additionally applies the one-line Transform_H fix and
# initialises Add_result (F4); "literal" compiles the decoder exactly as committed.
#
# usage: build_ref_gpu.sh NAME J L Z HFILE F MAXIT LEAST_TEST_FRAMES SNRTYPE [literal|fixed]
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
REF="${REF_ROOT:-/root/reference}/bldpc_实习"
[ -d "$REF" ] || { echo "reference tree not present ($REF) — keeping prebuilt baseline/_ref" >&2; exit 0; }
HDIR="${REF_HFILE_DIR:-/root/repo/cuda_ldpc_b200/data/bldpc}"
name=$1; J=$2; L=$3; Z=$4; hfile=$5; F=$6; maxit=$7; least=$8; snrtype=$9; variant=${10:-literal}
out="$here/_ref"; mkdir -p "$out"; tmp="$(mktemp -d)"; trap 'rm -rf "$tmp"' EXIT
NVCC=/usr/local/cuda/bin/nvcc
defs="-DREF_J=$J -DREF_L=$L -DREF_Z=$Z -DREF_F=$F -DREF_MAXIT=$maxit -DREF_LEAST_TEST_FRAMES=$least -DREF_SNRTYPE=$snrtype"
flags="-O3 -w -gencode arch=compute_100,code=sm_100 -ccbin /usr/bin/g++ -include $here/ref_gpu/define_override_gpu.cuh -I $REF $defs"
fix='s/__NOFIX__//'; init='s/__NOFIX__//'
if [ "$variant" = fixed ]; then
  fix='s/ - Z : index3;/ - Z : (Z - H[index1 * L + index0]) % Z + index3;/'
  init='s/float Add_result;/float Add_result = 0;/'
fi
for src in Simulation LDPC_Decoder LDPC_Encoder; do
  tr -d '\r' < "$REF/$src.cu" \
   | sed "s|char file\[100\] = \"PON_LDPC.txt\";|char file[200] = \"$HDIR/$hfile\";|" \
   | sed "$init" | sed "$fix" > "$tmp/$src.cu"
  $NVCC $flags -dc -o "$tmp/$src.o" "$tmp/$src.cu"
done
$NVCC $flags -dc -o "$tmp/harness.o" "$here/ref_gpu/ref_gpu_harness.cu"
$NVCC -gencode arch=compute_100,code=sm_100 -ccbin /usr/bin/g++ -o "$out/bldpc_gpu_$name" "$tmp"/*.o
echo "built $out/bldpc_gpu_$name"
