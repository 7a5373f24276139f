#!/bin/bash
# Round-1 profiling pass (run under gpurun): plain run first, then the ncu passes of B200_PROFILING.md
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench_pair100k.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_ncu1.log 2>&1
python tools/profile_one.py 100000 x 3 > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:psa_pair16 -s 1 -c 1 -f -o gpurun_out/r01_psa_pair16_100k_scoreonly \
    python tools/profile_one.py 100000 x 3 > gpurun_out/prof_ncu2.log 2>&1
python tools/profile_one.py 100000 trace 3 > gpurun_out/prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"psa_pair16|psa_trace" -s 2 -c 2 -f -o gpurun_out/r01_psa_pair16_100k_traced \
    python tools/profile_one.py 100000 trace 3 > gpurun_out/prof_ncu3.log 2>&1
python tools/batch_time.py 200000 250 x -4 > gpurun_out/prof_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:psa_batch16 -s 1 -c 1 -f -o gpurun_out/r01_psa_batch16_250bp \
    python tools/batch_time.py 200000 250 x -4 > gpurun_out/prof_ncu4.log 2>&1
tail -n 2 gpurun_out/prof_ncu*.log
