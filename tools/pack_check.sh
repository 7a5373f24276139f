#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_psa_packed_gpu.py tests/test_psa_gpu.py tests/test_psa_batch_gpu.py -x -q 2>&1 | tail -30 ) > gpurun_out/pack_pytest.log
( echo "score-only:"; timeout 120 python tools/profile_one.py 100000 x 4
  timeout 120 python tools/profile_one.py 200000 x 4
  echo "traced:"; for c in 4 8; do TSTA_PSA_C16=$c timeout 120 python tools/profile_one.py 100000 trace 4; done
  TSTA_PSA_CH=16 timeout 120 python tools/profile_one.py 100000 trace 4
  timeout 120 python tools/profile_one.py 200000 trace 4
  echo "int32 traced:"; TSTA_PSA_PACKED=0 timeout 120 python tools/profile_one.py 100000 trace 4 ) > gpurun_out/pack_pair.log 2>&1
( timeout 300 python tools/batch_time.py 1000000 250 x -4; timeout 300 python tools/batch_time.py 10000 10000 x -4
  timeout 300 python tools/batch_time.py 1000000 250 trace -4; timeout 300 python tools/batch_time.py 2000 10000 trace -4 ) > gpurun_out/pack_batch.log 2>&1
tools/pipe_rates.bin > gpurun_out/pipe_rates.log 2>&1
echo done
