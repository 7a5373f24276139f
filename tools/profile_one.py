"""Run one pairwise configuration a couple of times (for ncu / quick timing).
usage: python tools/profile_one.py N [trace] [reps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tsta_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
trace = len(sys.argv) > 2 and sys.argv[2] == "trace"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
import os
m = int(os.environ.get("M_ROWS", n))
x, y = synth.read_pair(n, m, int(os.environ.get("SEED", 1)))
p = api.PsaPlan(trace)
p.upload(x, y)
for _ in range(reps):
    f, t = p.run()
print("m=%d " % m, end="")
print("n=%d trace=%s fill %.3f ms trace %.3f ms %.1f GCUPS" % (n, trace, f, t, n * m / f / 1e6), p.info(), p.trace_stats() if trace else "")
