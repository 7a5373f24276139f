"""Chain graphs for the POA timing tools."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from tsta_b200 import api, synth

def chain(nrows, read, two=0.0):
    rp, pr = [0], []
    rng = np.random.default_rng(1)
    for r in range(nrows):
        ps = [-1] if r == 0 else [r - 1]
        if r > 3 and rng.random() < two: ps.append(r - 3)
        pr += ps; rp.append(len(pr))
    base = np.frombuffer((read * (nrows // len(read) + 1))[:nrows], dtype=np.uint8)
    hneg1 = (-6 - 2 * np.arange(nrows)).astype(np.int32)
    return np.array(rp, np.int32), np.array(pr, np.int32), base, hneg1

