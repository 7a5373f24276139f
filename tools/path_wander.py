"""How far the optimal path of a synthetic pair strays from the main diagonal (sizes the traceback corridor).
usage: python tools/path_wander.py n [n ...]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from tsta_b200 import api, synth

for n in map(int, sys.argv[1:]):
    x, y = synth.random_read(n, 1), synth.random_read(n, 2)
    plan = api.PsaPlan(True)
    plan.upload(x, y)
    plan.run()
    r = plan.fetch()
    st = plan.trace_stats()
    plan.close()
    A = np.frombuffer(r.a, dtype=np.uint8) != 45
    B = np.frombuffer(r.b, dtype=np.uint8) != 45
    d = np.cumsum(A) - np.cumsum(B)
    print("n=%d: alignment %d columns, j - i in [%d, %d], n^(2/3) = %.0f, corridors %d, tiles %d of %d"
          % (n, len(r.a), d.min(), d.max(), n ** (2 / 3), st["corridors"], st["tiles"], st["tiles_of_matrix"]))
