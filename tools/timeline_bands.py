"""Per-band durations and SM placement of a single pair's fill (TSTA_PSA_TIMELINE=1 stamps).
usage: python tools/timeline_bands.py [n]"""
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
x, y = synth.read_pair(n, n, 1)
p = api.PsaPlan(False)
p.upload(x, y)
for _ in range(3):
    f, t = p.run()
nb = p.info()["bands"]
tl = np.zeros((nb, 16), dtype=np.uint64)
L = api.load()
L.tsta_psa_plan_timeline.argtypes = [C.c_void_p, C.c_void_p]
api._check(L.tsta_psa_plan_timeline(p._h, tl.ctypes.data))
tl = tl.astype(np.int64)
t0 = tl[:, 0].min()
start, end, smid = tl[:, 1] - t0, tl[:, 7] - t0, tl[:, 11]
start[0] = tl[0, 0] - t0
dur = end - start
print("fill %.3f ms; bands %d" % (f, nb))
print("band  sm  start_us  dur_us  cyc/step   co-resident bands")
by_sm = {}
for b in range(nb):
    by_sm.setdefault(int(smid[b]), []).append(b)
for b in list(range(0, 24)) + list(range(24, nb, 16)):
    print("%4d %3d %9.1f %8.1f %7.1f   %s" % (b, smid[b], start[b] / 1e3, dur[b] / 1e3, dur[b] * 1.965 / n, by_sm[int(smid[b])]))
print("bands per SM: min %d max %d; SMs used %d" % (min(map(len, by_sm.values())), max(map(len, by_sm.values())), len(by_sm)))
d = dur * 1.965 / n
print("cycles/step by number of bands on the SM:")
for k in sorted(set(map(len, by_sm.values()))):
    sel = [b for v in by_sm.values() if len(v) == k for b in v]
    print("  %d bands/SM: %d bands, mean %.1f cycles/step" % (k, len(sel), d[sel].mean()))
