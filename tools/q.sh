for c in 3 4; do echo "C16=$c"; TSTA_PSA_C16=$c timeout 120 python tools/profile_one.py 100000 x 4; done
TSTA_PSA_C16=6 timeout 120 python tools/profile_one.py 200000 x 4
