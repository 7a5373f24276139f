timeout 300 python tools/batch_time.py 2000 10000 trace -4 | tail -n 2
timeout 300 python tools/batch_time.py 3000 10000 trace -4 | tail -n 1
