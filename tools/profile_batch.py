"""Run one batched configuration a couple of times (for ncu / quick timing).
usage: python tools/profile_batch.py NPAIRS LEN [trace] [reps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
from tsta_b200 import api

npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
ln = int(sys.argv[2]) if len(sys.argv) > 2 else 250
trace = len(sys.argv) > 3 and sys.argv[3] == "trace"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
import numpy as np
a, b = bench.make_batch(npairs, ln, 0)
s0, s1 = a.reshape(-1), b.reshape(-1)
o0 = o1 = np.arange(npairs, dtype=np.int64) * ln
l0 = l1 = np.full(npairs, ln, dtype=np.int32)
for _ in range(reps):
    r = api.psa_batch_arrays(s0, o0, l0, s1, o1, l1, trace=trace, packed=trace)
print("pairs %d x %d bp trace=%s kernels %.3f ms %.1f GCUPS" % (npairs, ln, trace, r.kernel_ms, float((l0.astype('int64') * l1).sum()) / r.kernel_ms / 1e6))
