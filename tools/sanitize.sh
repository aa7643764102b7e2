#!/bin/bash
# compute-sanitizer over every kernel family of libldpc_b200.so on small batches (run on the GPU box):
#   tools/sanitize.sh [out_dir]   -> out_dir/sanitize_{racecheck,memcheck}_{c1,c2,c3,nb}.log + a one-line verdict each
# racecheck = shared-memory hazards (the two round-1 races lived in the syndrome pass of the layered int8 kernel),
# memcheck = out-of-bounds / misaligned global + shared accesses.  SURVEY section 5 lists both as the B200 build's job.
# NOTE (round 2): compute-sanitizer is CLOSED on the graft GPU pool (every invocation prints "compute-sanitizer is closed
# on this pool" and exits 86); there the substitute is tests/test_stress_gpu.py (profiles/r02_stress_canary.txt).
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
if $CS --version 2>&1 | grep -q "closed on this pool"; then echo "compute-sanitizer is closed on this pool: use tests/test_stress_gpu.py"; exit 86; fi
rc_all=0
for tool in racecheck memcheck; do
  for which in c1 c3 c2 nb; do
    log="$OUT/sanitize_${tool}_${which}.log"
    extra=""
    [ "$tool" = racecheck ] && extra="--racecheck-report all"
    timeout 900 $CS --tool $tool $extra --print-limit 20 python tools/sanitize_cases.py $which > "$log" 2>&1
    rc=$?
    verdict=$(grep -E "RACECHECK SUMMARY|ERROR SUMMARY" "$log" | tail -1)
    echo "$tool $which rc=$rc ${verdict}"
    [ $rc -ne 0 ] && rc_all=1
    echo "$verdict" | grep -qE " 0 hazards displayed \(0 errors, 0 warnings\)|ERROR SUMMARY: 0 errors" || rc_all=1
  done
done
exit $rc_all
