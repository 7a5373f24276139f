"""Short form of tools/poa_model.py: cycles per row of the POA fill on chain graphs (TSTA_POA_KERNEL=skew|rows, TSTA_POA_C=2|4)."""
import sys
sys.path.insert(0, "/root/repo")
sys.path.insert(0, "/root/repo/tools")
from poa_model_lib import chain
from tsta_b200 import api, synth

poa = api.Poa()
for ln, nrows in ((128, 20000), (2560, 20000), (50000, 20000), (50000, 60000)):
    read = synth.random_read(ln, 3)
    g = chain(nrows, read, 0.2)
    poa.fill(*g, read)
    ms = min(poa.fill(*g, read)[1] for _ in range(3))
    print("len %6d rows %6d: %.3f ms = %.0f cycles per row" % (ln, nrows, ms, ms * 1.965e6 / nrows), flush=True)
