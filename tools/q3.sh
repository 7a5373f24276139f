timeout 600 python -m pytest tests/test_psa_gpu.py -x -q 2>&1 | tail -3
TSTA_TRACE_DEBUG=1 timeout 120 python tools/profile_one.py 100000 trace 2
