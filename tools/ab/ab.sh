#!/usr/bin/env bash
# A/B builds of the library in alternating fresh processes (developer tool)
cp cuda_ldpc_b200/libldpc_b200.so /tmp/lib_keep.so
for v in "$@"; do
  cp tools/ab/lib$v.so cuda_ldpc_b200/libldpc_b200.so
  echo "== $v"; python tools/ab/one.py
done
cp /tmp/lib_keep.so cuda_ldpc_b200/libldpc_b200.so
