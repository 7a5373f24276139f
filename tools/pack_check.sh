#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_psa_packed_gpu.py tests/test_psa_gpu.py tests/test_psa_batch_gpu.py -x -q 2>&1 | tail -30 ) > gpurun_out/pack_pytest.log
( for c in 2 3 4; do for ch in 16 32; do echo "C16=$c CH=$ch"; TSTA_PSA_C16=$c TSTA_PSA_CH=$ch timeout 120 python tools/profile_one.py 100000 x 4; done; done
  echo "200k:"; for c in 6 8; do TSTA_PSA_C16=$c timeout 120 python tools/profile_one.py 200000 x 4; done
  echo "20k:"; timeout 120 python tools/profile_one.py 20000 x 4
  echo "int32:"; TSTA_PSA_PACKED=0 timeout 120 python tools/profile_one.py 100000 x 4 ) > gpurun_out/pack_pair.log 2>&1
( timeout 300 python tools/batch_time.py 1000000 250 x -4; timeout 300 python tools/batch_time.py 10000 10000 x -4 ) > gpurun_out/pack_batch.log 2>&1
echo done
