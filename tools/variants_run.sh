#!/bin/bash
# experiment: the same timeline over several -D builds of the library (tools/timeline.py)
for v in "" "$@"; do
  if [ -z "$v" ]; then unset TSTA_LIB; else export TSTA_LIB=$PWD/tsta_b200/_build/libtsta_cuda_$v.so; fi
  echo "=== variant ${v:-default}"
  timeout 120 python tools/timeline.py 100000 | grep -v "^band"
done
