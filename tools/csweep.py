import os, sys
sys.path.insert(0, '/root/repo')
from tsta_b200 import api, synth
for n in (100000, 200000):
    x, y = synth.read_pair(n, n, 1)
    for tr, cs in ((False, (2, 3, 4, 6)), (True, (2, 4, 8))):
        for c in cs:
            os.environ["TSTA_PSA_C16"] = str(c)
            p = api.PsaPlan(tr); p.upload(x, y)
            for _ in range(2): p.run()
            ts = [p.run() for _ in range(4)]
            f = min(t[0] for t in ts); t = min(t[1] for t in ts)
            print(n, tr, c, "fill ms %.3f trace ms %.3f GCUPS fill %.1f" % (f, t, n * n / f / 1e6), p.info(), flush=True)
            p.close()
