#!/usr/bin/env bash
# every reference-GPU baseline binary the round-2 measurements use (SURVEY Appendix A constants; F = 4096 as committed)
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
B="$here/build_ref_gpu.sh"
#        name        J  L  Z    H file                     F    maxIT least snrtype variant
# timing binaries: "fixed" = Add_result initialised + the one-line Transform_H fix (F3); kernels, launch
# shapes, memory traffic and host loop are the reference's.  The literal build at F = 4096 is kept to show what the
# committed code does on sm_100: its uninitialised accumulator makes every hard decision 0, so the genie test
# "decodes" every frame after 2-3 iterations whatever the noise (profiles/r02_reference_gpu.txt).
$B C1_time          4 24   96 J4_L24_Z96_BlockH.txt     4096  10   4096  1 fixed
$B C2_time         15 30 1280 J15_L30_Z1280_BlockH.txt  4096  10   4096  0 fixed
$B C3_time         12 69  256 PON_LDPC.txt              4096  10   4096  1 fixed
$B C1_time_literal  4 24   96 J4_L24_Z96_BlockH.txt     4096  10   4096  1 literal
$B C1_fer_literal   4 24   96 J4_L24_Z96_BlockH.txt      256  10   1024  1 literal
$B C1_fer_fixed     4 24   96 J4_L24_Z96_BlockH.txt      256  10   1024  1 fixed
