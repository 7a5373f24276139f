nvidia-smi -L
timeout 600 python -m pytest tests/test_psa_multigpu.py tests/test_psa_packed_gpu.py -x -q 2>&1 | tail -8
timeout 120 python tools/profile_one.py 100000 x 4
