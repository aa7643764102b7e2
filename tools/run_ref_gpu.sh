#!/bin/bash
# Runs the reference's own GPU decoder (baseline/_ref, built by baseline/build_all.sh) on the GPU box:
#   tools/run_ref_gpu.sh [out_file]
out=${1:-gpurun_out/r02_reference_gpu.jsonl}
: > "$out"
R=baseline/_ref
for spec in "C1_time 3.0 5" "C2_time 2.0 3" "C3_time 3.0 3" "C3_time 4.5 3" "C1_time_literal 3.0 3"; do
  set -- $spec
  timeout 600 $R/bldpc_gpu_$1 time $2 $3 2>&1 | grep '^{' | sed "s/^{/{\"binary\": \"$1\", /" >> "$out"
done
timeout 900 $R/bldpc_gpu_C1_fer_literal fer 3.0 5.0 7.0 2>&1 | grep '^{' | sed 's/^{/{"binary": "C1_fer_literal", /' >> "$out"
timeout 900 $R/bldpc_gpu_C1_fer_fixed fer 2.5 3.0 3.5 2>&1 | grep '^{' | sed 's/^{/{"binary": "C1_fer_fixed", /' >> "$out"
cat "$out"
