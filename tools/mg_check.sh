timeout 900 python -m pytest tests/test_psa_multigpu.py -x -q 2>&1 | tail -6
