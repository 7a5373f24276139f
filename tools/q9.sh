TSTA_MSA_TIMING=1 timeout 900 python tools/msa_time.py 16 50000
