import sys, time
sys.path.insert(0,'/root/repo')
from tsta_b200 import api, synth
print(api.measure_dpx_peak())
for n in (20000, 100000):
    x,y = synth.read_pair(n,n,1)
    for tr in (False, True):
        p = api.PsaPlan(tr); p.upload(x,y)
        for _ in range(2): p.run()
        ts=[p.run() for _ in range(5)]
        f=min(t[0] for t in ts); t=min(t[1] for t in ts)
        print(n, tr, "fill ms %.3f trace ms %.3f GCUPS %.1f" % (f,t,n*n/f/1e6), p.info())
        p.close()
