#!/usr/bin/env bash
# A/B builds of the library with a caller-chosen probe script: ab2.sh probe.py libA libB ...
cp cuda_ldpc_b200/libldpc_b200.so /tmp/lib_keep.so
P=$1; shift
for v in "$@"; do cp tools/ab/lib$v.so cuda_ldpc_b200/libldpc_b200.so; echo "== $v"; python $P; done
cp /tmp/lib_keep.so cuda_ldpc_b200/libldpc_b200.so
