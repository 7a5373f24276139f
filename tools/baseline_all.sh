#!/bin/bash
# One GPU-box pass over every workload shape; logs to gpurun_out/baseline_*.log
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/baseline_gpu.log 2>&1
nproc >> gpurun_out/baseline_gpu.log; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/baseline_gpu.log; free -g >> gpurun_out/baseline_gpu.log
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/baseline_pytest.log
tools/pipe_rates.bin > gpurun_out/baseline_pipe_rates.log 2>&1
( timeout 300 python tools/quick_time.py ) > gpurun_out/baseline_pair.log 2>&1
( timeout 300 python tools/profile_one.py 200000 trace 3 ) >> gpurun_out/baseline_pair.log 2>&1
( timeout 300 python tools/profile_one.py 200000 x 3 ) >> gpurun_out/baseline_pair.log 2>&1
( for O in -4 0; do for tr in x trace; do
    timeout 300 python tools/batch_time.py 1000000 250 $tr $O
    timeout 300 python tools/batch_time.py 10000 10000 $tr $O
  done; done ) > gpurun_out/baseline_batch.log 2>&1
( TSTA_MSA_TIMING=1 timeout 600 python tools/msa_time.py 16 10000
  TSTA_MSA_TIMING=1 timeout 900 python tools/msa_time.py 16 50000 ) > gpurun_out/baseline_msa.log 2>&1
( timeout 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/baseline_bench.json 2> gpurun_out/baseline_bench.err
echo done
