#!/bin/bash
# ncu --set full of the traced 250 bp batch kernels (200k pairs), reports into gpurun_out/
mkdir -p gpurun_out
python tools/profile_batch.py 200000 250 trace 2 > gpurun_out/pb.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"psa_batch16_dual_kernel|psa_batch_trace_kernel" --launch-skip 2 --launch-count 2 \
    -o gpurun_out/r02g_batch250_trace -f python tools/profile_batch.py 200000 250 trace 2 > gpurun_out/ncu_pb.log 2>&1
python tools/profile_batch.py 10000 10000 "" 1 >> gpurun_out/pb.log 2>&1
