timeout 120 python tools/profile_one.py 100000 trace 4
timeout 120 python tools/profile_one.py 200000 trace 3
timeout 300 python tools/batch_time.py 1000000 250 trace -4 | tail -n 1
