import sys, os
sys.path.insert(0, "/root/repo")
from tsta_b200 import api, synth
x, y = synth.random_read(100000, 1), synth.random_read(100000, 2)
for tr in (False, True):
    p = api.PsaPlan(tr)
    p.upload(x, y)
    ts = []; tt=[]
    for k in range(12):
        f, t = p.run()
        if k >= 4: ts.append(f); tt.append(t)
    print("trace" if tr else "score", "fill ms min %.3f mean %.3f trace-chain min %.3f" % (min(ts), sum(ts) / len(ts), min(tt)), p.info())
    p.close()
