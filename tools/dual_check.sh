timeout 900 python -m pytest tests/test_psa_packed_gpu.py tests/test_psa_batch_gpu.py tests/test_cli.py -x -q 2>&1 | tail -6
timeout 300 python tools/batch_time.py 1000000 250 trace -4 | tail -n 1
TSTA_PSA_DUAL=0 timeout 300 python tools/batch_time.py 1000000 250 trace -4 | tail -n 1
timeout 300 python tools/batch_time.py 1000000 100 trace -4 | tail -n 1
