"""Time tsta_psa_batch on synthetic batches (BASELINE config 4 shapes, scalable).
usage: python tools/batch_time.py NPAIRS LEN [trace] [O]"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tsta_b200 import api, synth


def make_batch(npairs: int, ln: int, seed: int = 5):
    """Half independent pairs, half 5 %-divergent copies, all of length `ln` (s1 cut/padded to <= ln)."""
    a = np.frombuffer(synth.random_read(npairs * ln, seed), dtype=np.uint8).reshape(npairs, ln)
    b = np.frombuffer(synth.random_read(npairs * ln, seed + 1), dtype=np.uint8).reshape(npairs, ln).copy()
    # divergent copies for odd pairs: substitutions at 5 % (keeps equal lengths so the batch stays rectangular)
    rnd = np.random.default_rng(seed)
    odd = np.arange(npairs) % 2 == 1
    sub = rnd.random((npairs, ln)) < 0.05
    b[odd] = np.where(sub[odd], b[odd], a[odd])
    off = np.arange(npairs, dtype=np.int64) * ln
    lens = np.full(npairs, ln, dtype=np.int32)
    return a.reshape(-1).copy(), off, lens, b.reshape(-1).copy(), off.copy(), lens.copy()


if __name__ == "__main__":
    npairs, ln = int(sys.argv[1]), int(sys.argv[2])
    trace = len(sys.argv) > 3 and sys.argv[3] == "trace"
    O = int(sys.argv[4]) if len(sys.argv) > 4 else -4
    s0, o0, l0, s1, o1, l1 = make_batch(npairs, ln)
    for rep in range(3):
        t0 = time.perf_counter()
        r = api.psa_batch_arrays(s0, o0, l0, s1, o1, l1, O=O, trace=trace)
        dt = time.perf_counter() - t0
        cells = float(npairs) * ln * ln
        print("npairs=%d len=%d trace=%s O=%d: kernel %.3f ms (%.1f GCUPS)  call %.3f s (%.1f GCUPS)  sum(max)=%d sum(corner)=%d"
              % (npairs, ln, trace, O, r.kernel_ms, cells / r.kernel_ms / 1e6, dt, cells / dt / 1e9,
                 int(r.maxsorce.astype(np.int64).sum()), int(r.corner.astype(np.int64).sum())), flush=True)
