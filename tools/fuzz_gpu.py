"""One-off randomized parity sweep on the GPU with seeds the test-suite does not use:
POA fill + trace + walk on random DAGs (incl. 'N' reads, many heads, wide in-degree, far
predecessors), single pairs and batches against the oracle.  usage: python tools/fuzz_gpu.py [seed0] [n]"""
import random, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from oracle import binding as ob
from tsta_b200 import api, synth
import test_poa_gpu as tp

seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 60
PRM = [(2, -5, -2, -4), (2, -3, -2, -4), (1, -1, -1, 0), (2, -5, -2, 0), (1, -3, -1, -5), (5, -4, -1, -10), (3, -2, -3, -1)]
bad = 0
t0 = time.time()
poa = api.Poa()
for k in range(count):
    rnd = random.Random(seed0 + k)
    M, X, E, O = rnd.choice(PRM)
    nrows = rnd.choice([1, 2, 33, 64, 65, 200, 511, 1200, 2500])
    ln = rnd.choice([1, 2, 4, 5, 127, 128, 129, 255, 256, 257, 640, 1000])
    L = rnd.choice([64, 128, 640])
    read = bytearray(synth.random_read(ln, seed0 * 7 + k))
    if rnd.random() < 0.4:
        for i in range(ln):
            if rnd.random() < 0.05: read[i] = ord("N")
    read = bytes(read)
    g = tp.random_dag(nrows, rnd, E, O, far=rnd.choice([0.0, 0.03, 0.3]), heads=rnd.choice([0.0, 0.01, 0.08]),
                      maxin=rnd.choice([2, 4, 5, 9, 20]), backbone=read.replace(b"N", b"C"))
    src, es, fs, last = ob.poa_fill(*g, read, M, X, E, O, L)
    glast, _ = poa.fill(*g, read, M, X, E, O, L)
    ok = (glast == last).all()
    want, mask = tp.expected_trace(src, es, fs, g[0], g[1])
    ok = ok and ((poa.trace() & mask) == (want & mask)).all()
    for start in {nrows - 1, rnd.randrange(nrows)}:
        w = ob.poa_walk(g[0], g[1], src, es, fs, start)
        got, _ = poa.walk(start)
        ok = ok and len(got) == len(w) and all((got[f] == w[f]).all() for f in ("row", "col", "kind", "pred"))
    if not ok:
        bad += 1
        print("POA MISMATCH seed", seed0 + k, (M, X, E, O), nrows, ln, L, flush=True)
poa.close()
print("poa cases %d bad %d (%.1f s)" % (count, bad, time.time() - t0), flush=True)

t0 = time.time()
lens = [1, 2, 3, 31, 32, 33, 63, 64, 65, 191, 192, 193, 255, 256, 257, 300, 767, 768, 769, 1535, 1536, 1537, 2000, 4100]
for k in range(count):
    rnd = random.Random(seed0 + 100000 + k)
    prm = rnd.choice(PRM)
    x, y = synth.read_pair(rnd.choice(lens), rnd.choice(lens), seed0 + 31 * k, rnd.choice([None, 0.02, 0.1, 0.3]))
    y = y or b"G"
    o = ob.psa(x, y, *prm)
    r = api.psa(x, y, *prm, trace=True)
    r0 = api.psa(x, y, *prm, trace=False)
    if (r.maxsorce, r.corner) != (o.maxsorce, o.corner) or (r0.maxsorce, r0.corner) != (o.maxsorce, o.corner) or r.output_bytes() != o.output_bytes():
        bad += 1
        print("PSA MISMATCH seed", seed0 + k, prm, len(x), len(y), flush=True)
print("psa cases %d, total bad %d (%.1f s)" % (count, bad, time.time() - t0), flush=True)

t0 = time.time()
for k in range(4):
    rnd = random.Random(seed0 + 200000 + k)
    prm = rnd.choice(PRM[:5])
    pairs = []
    for q in range(400):
        x, y = synth.read_pair(rnd.choice(lens[:17]), rnd.choice(lens[:17]), seed0 + 977 * k + q, rnd.choice([None, 0.05, 0.2]))
        pairs.append((x, y or b"T"))
    import os
    os.environ["TSTA_BATCH_CHUNKS"] = str(rnd.choice([1, 2, 5, 8]))
    rb0 = api.psa_batch(pairs, *prm, trace=False)
    rb1 = api.psa_batch(pairs, *prm, trace=True)
    for q, (x, y) in enumerate(pairs):
        o = ob.psa(x, y, *prm)
        if (rb0.maxsorce[q], rb0.corner[q]) != (o.maxsorce, o.corner) or (rb1.maxsorce[q], rb1.corner[q]) != (o.maxsorce, o.corner) \
                or rb1.output_bytes(q) != o.output_bytes():
            bad += 1
            print("BATCH MISMATCH", seed0 + k, q, prm, len(x), len(y), flush=True)
            break
print("batch cases 4 x 400, total bad %d (%.1f s)" % (bad, time.time() - t0), flush=True)
sys.exit(1 if bad else 0)
