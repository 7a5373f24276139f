"""POA fill kernel time against rows and bands on chain graphs: slope = cycles per step, intercept = ramp."""
import sys
sys.path.insert(0, "/root/repo")
sys.path.insert(0, "/root/repo/tools")
from poa_model_lib import chain
from tsta_b200 import api, synth

poa = api.Poa()
for ln in (128, 2560, 20000, 50000):
    read = synth.random_read(ln, 3)
    for nrows in (10000, 20000, 40000):
        for two in (0.0, 0.25):
            g = chain(nrows, read, two)
            poa.fill(*g, read)
            ms = min(poa.fill(*g, read)[1] for _ in range(3))
            print("len %6d bands %4d rows %6d two-pred %.2f: %.3f ms = %.0f cycles per row" % (ln, (ln + 127) // 128, nrows, two, ms, ms * 1.965e6 / nrows), flush=True)
