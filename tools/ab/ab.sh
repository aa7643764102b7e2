#!/usr/bin/env bash
# A/B two builds of the library in alternating fresh processes (developer tool)
for rep in 1 2 3; do
  for v in A_v4 B_v7; do
    cp tools/ab/lib$v.so cuda_ldpc_b200/libldpc_b200.so
    echo "== $v rep $rep"; python tools/ab/one.py
  done
done
