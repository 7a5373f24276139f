"""Where a single pair's time goes, band by band (TSTA_PSA_TIMELINE=1 stamps).
usage: [TSTA_LIB=...] python tools/timeline.py [n] [trace]"""
import ctypes as C
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ["TSTA_PSA_TIMELINE"] = "1"
if not os.environ.get("TSTA_LIB"):      # the stamps are compiled in only with -DTSTA_TIMELINE=1 (build it where nvcc is: python tools/timeline.py --build)
    os.environ["TSTA_LIB"] = str(Path(__file__).resolve().parents[1] / "tsta_b200" / "_build" / "libtsta_cuda_timeline.so")
if "--build" in sys.argv:
    from tsta_b200 import build
    print(build.build_variant("timeline", ["TSTA_TIMELINE=1"]))
    sys.exit(0)
import numpy as np
from tsta_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
trace = len(sys.argv) > 2 and sys.argv[2] == "trace"
x, y = synth.read_pair(n, n, 1)
p = api.PsaPlan(trace)
p.upload(x, y)
for _ in range(3):
    f, t = p.run()
info = p.info()
nb = info["bands"]
tl = np.zeros((nb, 16), dtype=np.uint64)
L = api.load()
L.tsta_psa_plan_timeline.argtypes = [C.c_void_p, C.c_void_p]
api._check(L.tsta_psa_plan_timeline(p._h, tl.ctypes.data))
tl = tl.astype(np.int64)
stats = tl[:, 8:11].copy()
tl = tl[:, :8]
t0 = tl[:, 0].min()
tl -= t0
print("lib %s n=%d trace=%s: fill %.3f ms, trace %.3f ms, %s" % (os.environ.get("TSTA_LIB", "default"), n, trace, f, t, info))
names = ["entry", "first rows in", "chunk1", "chunk2", "chunk3", "chunk4", "chunk5", "end"]
for b in (0, 1, 2, 10, 100, nb // 2, nb - 2, nb - 1):
    print("band %4d: " % b + "  ".join("%s %8.1f us" % (names[k], tl[b, k] / 1e3) for k in range(8)))
d = np.diff(tl[:, 1])[1:]            # band b's first rows arrive this long after band b-1's
print("first-rows-in spacing between consecutive bands: mean %.2f us  median %.2f  p90 %.2f  (x %d bands = %.3f ms)"
      % (d.mean() / 1e3, np.median(d) / 1e3, np.percentile(d, 90) / 1e3, nb - 2, d.sum() / 1e6))
for k in range(1, 6):
    c = tl[1:-1, k + 1] - tl[1:-1, k]
    print("  duration of chunk %d (32 steps): mean %.2f us  median %.2f us" % (k - 1, c.mean() / 1e3, np.median(c) / 1e3))
body = tl[1:-1, 7] - tl[1:-1, 6]
steps = (n + 63 + 31) // 32 - 5
print("  steady state: %.3f us per 32-step chunk (%.1f cycles/step at 1.965 GHz)" % (body.mean() / 1e3 / steps, body.mean() * 1.965 / steps / 32))
print("  band 0 entry->end %.3f ms; last band end at %.3f ms" % ((tl[0, 7] - tl[0, 0]) / 1e6, tl[-1, 7] / 1e6))
imp = stats[1:-1]
print("  imports per band %.0f, stale polls per band %.1f (%.3f per import), cycles per import %.0f"
      % (imp[:, 2].mean(), imp[:, 0].mean(), imp[:, 0].sum() / max(imp[:, 2].sum(), 1), imp[:, 1].sum() / max(imp[:, 2].sum(), 1)))
