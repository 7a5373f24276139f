timeout 900 python -m pytest tests/test_poa_gpu.py -x -q 2>&1 | tail -2
TSTA_MSA_TIMING=1 timeout 900 python tools/msa_time.py 16 50000 2>&1 | head -3
timeout 100 python tools/poa_model.py 2>/dev/null | head -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
