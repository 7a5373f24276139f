#!/bin/bash
# A/B of the streamed left boundary on the 100 kbp pair (tools/quick_pair.py): variants = names of _build/libtsta_cuda_<v>.so
mkdir -p gpurun_out
{
for v in "" "$@" ""; do
  if [ -z "$v" ]; then unset TSTA_LIB; else export TSTA_LIB=$PWD/tsta_b200/_build/libtsta_cuda_$v.so; fi
  timeout 120 python tools/quick_pair.py
done
} > gpurun_out/stream_ab.log 2>&1
