timeout 900 python -m pytest tests/test_poa_gpu.py -x -q 2>&1 | tail -3
TSTA_MSA_TIMING=1 timeout 900 python tools/msa_time.py 16 50000
