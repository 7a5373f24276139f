import sys, os
import numpy as np
sys.path.insert(0, "/root/repo")
from tsta_b200 import api, synth
x, y = synth.random_read(100000, 1), synth.random_read(100000, 2)
p = api.PsaPlan(False)
p.upload(x, y)
ts = []
for k in range(12):
    f, t = p.run()
    if k >= 4: ts.append(f)
r = p.fetch()
lr = p.lastrow()
print("maxsorce", r.maxsorce, "corner", r.corner, "lastrow crc", int((lr.astype("int64") * (1 + (np.arange(lr.size) % 1009))).sum()))
print(os.environ.get("TSTA_LIB", "default").split("/")[-1], "fill ms min %.3f mean %.3f" % (min(ts), sum(ts) / len(ts)), p.info())
