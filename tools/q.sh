timeout 900 python -m pytest tests/test_psa_gpu.py -x -q 2>&1 | tail -5
timeout 120 python tools/profile_one.py 100000 trace 4
timeout 120 python tools/profile_one.py 200000 trace 4
