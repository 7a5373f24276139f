"""One POA fill on a chain graph (tools/poa_one.py <len> <rows> [two-pred fraction]): the launch ncu captures."""
import sys
sys.path.insert(0, '/root/repo')
sys.path.insert(0, '/root/repo/tools')
from poa_model_lib import chain
from tsta_b200 import api, synth
ln, nrows = int(sys.argv[1]), int(sys.argv[2])
two = float(sys.argv[3]) if len(sys.argv) > 3 else 0.2
read = synth.random_read(ln, 3)
g = chain(nrows, read, two)
poa = api.Poa()
for _ in range(3):
    ms = poa.fill(*g, read)[1]
print("len %d rows %d: %.3f ms = %.0f cycles per row" % (ln, nrows, ms, ms * 1.965e6 / nrows))
