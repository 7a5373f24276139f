TSTA_TRACE_DEBUG=1 timeout 120 python tools/profile_one.py 100000 trace 2
