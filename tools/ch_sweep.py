import os, sys
sys.path.insert(0, '/root/repo')
from tsta_b200 import api, synth
for n in (100000, 200000):
    x, y = synth.read_pair(n, n, 1)
    for tr in (False, True):
        for ch in (16, 32):
            os.environ["TSTA_PSA_CH"] = str(ch)
            p = api.PsaPlan(tr); p.upload(x, y)
            for _ in range(2): p.run()
            ts = [p.run() for _ in range(4)]
            f = min(t[0] for t in ts)
            print(n, tr, "CH", ch, "fill ms %.3f GCUPS %.1f" % (f, n * n / f / 1e6), p.info(), flush=True)
            p.close()
