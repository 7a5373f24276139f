#!/bin/bash
# A/B of library builds on the 100 kbp pair (tools/quick_pair.py): arguments = names of _build/libtsta_cuda_<v>.so made with
# tsta_b200.build.build_variant (e.g. nostream, nopin); "" = the default build.  Log: gpurun_out/stream_ab.log
mkdir -p gpurun_out
{
for v in "" "$@" ""; do
  if [ -z "$v" ]; then unset TSTA_LIB; else export TSTA_LIB=$PWD/tsta_b200/_build/libtsta_cuda_$v.so; fi
  timeout 120 python tools/quick_pair.py
done
} > gpurun_out/stream_ab.log 2>&1
