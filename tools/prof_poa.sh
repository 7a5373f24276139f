#!/bin/bash
# ncu --set full of one late-round POA fill (16 x 20 kbp job), after a plain run of the same command
mkdir -p gpurun_out
TSTA_MSA_TIMING=1 timeout 300 python tools/msa_time.py 16 20000 > gpurun_out/pp_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:poa_fill -s 12 -c 1 -f -o gpurun_out/r01_poa_fill_16x20k_bytekeys \
    tsta_b200/_build/TSTA_msa -i /tmp/msa_time/in.fa -o /tmp/msa_time/o2.txt > gpurun_out/pp_ncu.log 2>&1
tail -n 3 gpurun_out/pp_plain.log gpurun_out/pp_ncu.log
