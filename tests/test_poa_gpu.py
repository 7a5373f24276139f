"""Parity of the CUDA POA path with the oracle (cell by cell on random DAGs) and, end to end
through the TSTA_msa driver, with the real reference's outputs."""
import hashlib
import json
import os
import random
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import binding as ob
from tsta_b200 import synth
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "tsta_b200" / "_build"


def random_dag(nrows, rnd, E, O, far=0.03, heads=0.01, maxin=4, backbone=None):
    """A graph in topological row order shaped like a POA graph: mostly a chain, some short
    and a few long-range extra predecessors, occasional extra heads."""
    row_ptr, pred, base, hneg1 = [0], [], [], []
    for r in range(nrows):
        if r == 0 or rnd.random() < heads:
            ps = [-1]
        else:
            ps = [r - 1] if rnd.random() < 0.9 else [max(0, r - rnd.randint(2, 4))]
            k = rnd.choice([0, 0, 0, 1, 1, 2, maxin - 1])
            for _ in range(k):
                d = rnd.randint(2, 6) if rnd.random() > far else rnd.randint(20, 200)
                q = max(0, r - d)
                if q not in ps:
                    ps.append(q)
            rnd.shuffle(ps)
        pred += ps
        row_ptr.append(len(pred))
        base.append(backbone[r % len(backbone)] if backbone and rnd.random() < 0.8 else rnd.choice(b"ACGT"))
        hneg1.append(O + E if ps[0] < 0 else max(hneg1[p] for p in ps) + E)
    return (np.array(row_ptr, np.int32), np.array(pred, np.int32), np.array(base, np.uint8), np.array(hneg1, np.int32))


def expected_trace(src, es, fs, row_ptr, pred):
    """The oracle's three matrices in the 2-byte cell format csrc/poa_fill.cuh documents, and the mask
    of the bits the reference's matrices determine:
      byte 0: type << 6 | 63 - j      (3 diagonal, 2 vertical, 1 horizontal with j field 63)
      byte 1: 63 - j*, bit 6 = NOT (be+E >= H+O+E), bit 7 = (Fg+E >= H+O+E), both at the cell itself.
    Bit 7 of (r,c) is what fsource says about the step c -> c+1 (c-t-simd.c:767-769), so the last column
    is unconstrained; bit 6 of (p,c) is what a successor whose best-E predecessor is p recorded about p
    (776-782), so only such cells are constrained -- the walker never looks at the others."""
    s = src.astype(np.int32) & 0x7f
    kind, j = s // 42, s % 42
    lo = np.where(kind == 3, 64, np.where(kind == 0, 128 | (63 - j), 192 | (63 - j)))
    e = es.astype(np.int32)
    eopen = np.abs(e) > 42
    jstar = np.abs(e) - np.where(eopen, 43, 1)
    f = fs.astype(np.int32)
    nrows, ln = s.shape
    gF = np.zeros_like(f, dtype=bool)
    gF[:, :-1] = (np.abs(f[:, 1:]) != 2) | (f[:, :-1] < 0)
    notgE = np.zeros((nrows, ln), dtype=bool)
    seen = np.zeros((nrows, ln), dtype=bool)
    cols = np.arange(ln)
    for r in range(nrows):
        ps = pred[row_ptr[r]:row_ptr[r + 1]]
        p = ps[jstar[r]]
        ok = p >= 0
        gE = ~eopen[r] | (e[np.maximum(p, 0), cols] < 0)
        notgE[p[ok], cols[ok]] = ~gE[ok]
        seen[p[ok], cols[ok]] = True
    want = lo | ((63 - jstar) << 8) | (notgE.astype(np.int32) << 14) | (gF.astype(np.int32) << 15)
    mask = np.full((nrows, ln), 0x3fff, dtype=np.int32) | (seen.astype(np.int32) << 14)
    mask[:, :-1] |= 0x8000
    return want.astype(np.uint16), mask.astype(np.uint16)


def assert_trace(poa, src, es, fs, g):
    want, mask = expected_trace(src, es, fs, g[0], g[1])
    np.testing.assert_array_equal(poa.trace() & mask, want & mask)


@pytest.mark.parametrize("seed", range(5))
def test_fill_and_walk_on_random_dags(gpu, seed):
    rnd = random.Random(seed)
    poa = gpu.Poa()
    for k in range(4):
        M, X, E, O = rnd.choice([(2, -5, -2, -4), (2, -3, -2, -4), (1, -1, -1, 0), (2, -5, -2, 0), (1, -3, -1, -5)])
        nrows = rnd.choice([1, 7, 40, 129, 300, 700])
        ln = rnd.choice([1, 3, 127, 128, 129, 300, 513])
        L = rnd.choice([64, 128, 640])
        read = synth.random_read(ln, 50 * seed + k)
        g = random_dag(nrows, rnd, E, O, backbone=read)
        src, es, fs, last = ob.poa_fill(*g, read, M, X, E, O, L)
        glast, _ = poa.fill(*g, read, M, X, E, O, L)
        np.testing.assert_array_equal(glast, last)
        assert_trace(poa, src, es, fs, g)
        for start in {nrows - 1, rnd.randrange(nrows)}:
            want = ob.poa_walk(g[0], g[1], src, es, fs, start)
            got, _ = poa.walk(start)
            assert len(got) == len(want)
            for fld in ("row", "col", "kind", "pred"):
                np.testing.assert_array_equal(got[fld], want[fld])
    poa.close()


def test_long_graph_far_predecessors_and_wide_in_degree(gpu):
    rnd = random.Random(99)
    read = synth.random_read(1000, 5)
    g = random_dag(3000, rnd, -2, -4, far=0.2, heads=0.003, maxin=12, backbone=read)
    src, es, fs, last = ob.poa_fill(*g, read, L=640)
    poa = gpu.Poa()
    glast, _ = poa.fill(*g, read, L=640)
    np.testing.assert_array_equal(glast, last)
    assert_trace(poa, src, es, fs, g)
    want = ob.poa_walk(g[0], g[1], src, es, fs, 2999)
    got, _ = poa.walk(2999)
    assert len(got) == len(want) and (got["row"] == want["row"]).all() and (got["kind"] == want["kind"]).all()
    poa.close()


@pytest.mark.parametrize("seed", range(3))
def test_quirk_rows_n_columns_and_late_heads(gpu, seed):
    """The two quirks the kernel keeps out of its straight-line step: reads with 'N'
    (Q3, c-t-simd.c:461,481) and head rows at or beyond row L whose carried F is re-based at column L
    (Q1, 228-233) -- many heads, small L, so that both meet inside one strip."""
    rnd = random.Random(1000 + seed)
    poa = gpu.Poa()
    for k in range(4):
        M, X, E, O = rnd.choice([(2, -5, -2, -4), (2, -3, -2, -4), (1, -1, -1, 0), (1, -3, -1, -5)])
        nrows, ln, L = rnd.choice([200, 500, 900]), rnd.choice([130, 300, 700]), rnd.choice([64, 128])
        read = bytearray(synth.random_read(ln, 77 * seed + k))
        for i in range(ln):
            if rnd.random() < 0.06:
                read[i] = ord("N")
        read = bytes(read)
        g = random_dag(nrows, rnd, E, O, heads=0.05, backbone=read.replace(b"N", b"A"))
        src, es, fs, last = ob.poa_fill(*g, read, M, X, E, O, L)
        glast, _ = poa.fill(*g, read, M, X, E, O, L)
        np.testing.assert_array_equal(glast, last)
        assert_trace(poa, src, es, fs, g)
        for start in {nrows - 1, rnd.randrange(nrows)}:
            want = ob.poa_walk(g[0], g[1], src, es, fs, start)
            got, _ = poa.walk(start)
            assert len(got) == len(want)
            for fld in ("row", "col", "kind", "pred"):
                np.testing.assert_array_equal(got[fld], want[fld])
    poa.close()


@pytest.mark.parametrize("kernel,C", [("rows", 4), ("rows", 3), ("rows", 2), ("skew", 4), ("skew", 2)])
def test_both_fill_kernels_late_heads_dense_graphs(gpu, monkeypatch, kernel, C):
    """The row-synchronous kernel (poa_rows.cuh: F by scan, register forwarding, serial pass for late heads in
    the band that owns column L) and the skewed kernel (poa_fill.cuh), each strip width, on reads WITHOUT 'N'
    -- the inputs both kernels accept -- with many heads, small L, wide and far predecessors."""
    monkeypatch.setenv("TSTA_POA_KERNEL", kernel)
    monkeypatch.setenv("TSTA_POA_C", str(C))
    rnd = random.Random(4242 + C)
    poa = gpu.Poa()
    for k in range(10):
        M, X, E, O = rnd.choice([(2, -5, -2, -4), (1, -1, -1, 0), (2, -5, -2, 0), (1, -3, -1, -5), (5, -4, 0, -3)])
        # (L anywhere inside a lane's strip: 70, 99, 130 are not multiples of any strip width)
        nrows, ln, L = rnd.choice([90, 400, 1100]), rnd.choice([5, 131, 260, 900]), rnd.choice([64, 128, 640, 70, 99, 130])
        read = synth.random_read(ln, 900 + 13 * k + C)
        g = random_dag(nrows, rnd, E, O, far=rnd.choice([0.0, 0.2]), heads=rnd.choice([0.0, 0.06]),
                       maxin=rnd.choice([3, 6, 11]), backbone=read)
        src, es, fs, last = ob.poa_fill(*g, read, M, X, E, O, L)
        glast, _ = poa.fill(*g, read, M, X, E, O, L)
        np.testing.assert_array_equal(glast, last)
        assert_trace(poa, src, es, fs, g)
        want = ob.poa_walk(g[0], g[1], src, es, fs, nrows - 1)
        got, _ = poa.walk(nrows - 1)
        assert len(got) == len(want)
        for fld in ("row", "col", "kind", "pred"):
            np.testing.assert_array_equal(got[fld], want[fld])
    poa.close()


@pytest.mark.parametrize("seed", range(4))
def test_resident_graph_tables_built_on_the_device(gpu, seed):
    """tsta_poa_graph_update + tsta_poa_fill_resident (node records in NODE-id space, only the permutation per read;
    descriptors, predecessor table, spill slots and H(r,-1) made on the device) against the oracle, across several
    "rounds" in which nodes are added, predecessor lists change and the order is reshuffled -- including a read
    with 'N' (skewed kernel, hneg1 supplied) and rows with far predecessors."""
    rnd = random.Random(3100 + seed)
    poa = gpu.Poa()
    M, X, E, O = rnd.choice([(2, -5, -2, -4), (1, -1, -1, 0), (1, -3, -1, -5)])
    L = rnd.choice([64, 128, 640])
    ids = {}                                   # row of the first round -> stable node id (shuffled)
    sent = {}
    for rnd_no in range(3):
        nrows = [150, 420, 900][rnd_no] + 37 * seed
        ln = rnd.choice([90, 260, 700])
        read = bytearray(synth.random_read(ln, 40 * seed + rnd_no))
        if rnd_no == 2 and seed % 2 == 0:
            for i in range(0, ln, 11):
                read[i] = ord("N")
        read = bytes(read)
        g = random_dag(nrows, rnd, E, O, far=0.1, heads=0.03, maxin=8, backbone=read.replace(b"N", b"A"))
        row_ptr, pred, base, hneg1 = g
        # stable ids: rows seen before keep theirs, new rows draw fresh ones; rows are "re-sorted" simply by being
        # a different graph over a superset of the ids -- all the device sees is (records, permutation)
        fresh = [i for i in range(2 * nrows + 50) if i not in ids.values()]
        rnd.shuffle(fresh)
        for r in range(nrows):
            if r not in ids:
                ids[r] = fresh.pop()
        nnodes = max(ids.values()) + 1
        changed = []
        for r in range(nrows):
            ps = [ids[p] for p in pred[row_ptr[r]:row_ptr[r + 1]] if p >= 0]
            rec = (int(base[r]), ps)
            if sent.get(ids[r]) != rec:
                sent[ids[r]] = rec
                changed.append((ids[r], rec[0], rec[1]))
        before = poa.h2d_bytes()
        poa.graph_update(nnodes, changed)
        sort = np.array([ids[r] for r in range(nrows)], np.int32)
        has_n = b"N" in read
        glast, _ = poa.fill_resident(sort, read, M, X, E, O, L, hneg1=hneg1 if has_n else None)
        assert poa.h2d_bytes() - before <= 16 * len([c for c in changed if len(c[2]) <= 2]) + 40 * len(changed) + 8 * nrows + ln
        src, es, fs, last = ob.poa_fill(*g, read, M, X, E, O, L)
        np.testing.assert_array_equal(glast, last)
        assert_trace(poa, src, es, fs, g)
        for start in {nrows - 1, rnd.randrange(nrows)}:
            want = ob.poa_walk(g[0], g[1], src, es, fs, start)
            got, _ = poa.walk(start)
            assert len(got) == len(want)
            for fld in ("row", "col", "kind", "pred"):
                np.testing.assert_array_equal(got[fld], want[fld])
    poa.close()


def test_msa_full_export_and_resident_graph_agree(gpu, tmp_path, monkeypatch):
    """The driver's two export modes write the same file; the resident mode copies a fraction of the bytes (a short job:
    the one-off upload of the first read's nodes weighs more than at 16 x 50 kbp, where tests/test_full_size_gpu.py
    asks for a factor of five)."""
    reads = [synth.random_read(6000, 77)]
    reads += [synth.mutate(reads[0], 0.06, 500 + k)[:6000] for k in range(7)]
    fa = tmp_path / "r.fa"
    ob.write_fasta(fa, reads)
    out, nbytes = {}, {}
    for mode in ("resident", "full"):
        env = dict(os.environ, TSTA_MSA_TIMING="1")
        if mode == "full":
            env["TSTA_MSA_FULL_EXPORT"] = "1"
        fo = tmp_path / (mode + ".txt")
        p = subprocess.run([str(BIN / "TSTA_msa"), "-i", str(fa), "-o", str(fo)], capture_output=True, timeout=900, env=env)
        assert p.returncode == 0, p.stderr
        out[mode] = fo.read_bytes()
        line = [l for l in p.stderr.decode().splitlines() if l.startswith("export:")][0]
        nbytes[mode] = int(line.split()[1])
    assert out["resident"] == out["full"]
    assert nbytes["full"] >= 3 * nbytes["resident"], nbytes


def run_msa(reads, prm, W, tmp_path):
    fa, fo = tmp_path / "r.fa", tmp_path / "o.txt"
    ob.write_fasta(fa, reads)
    p = subprocess.run([str(BIN / "TSTA_msa"), "-M", str(prm[0]), "-X", str(prm[1]), "-E", str(prm[2]), "-O", str(prm[3]),
                        "-W", str(W), "-T", "2", "-i", str(fa), "-o", str(fo)], capture_output=True, timeout=900)
    assert p.returncode == 0, p.stderr
    return fo.read_bytes(), p.stdout, p.stderr


def test_msa_driver_example_golden(gpu, tmp_path):
    g = json.load(open(GOLDEN / "msa_golden.json"))
    reads = [r.encode() for r in g["example_reads"]]
    for case in g["example"]:
        out, so, se = run_msa(reads, case["params"], case["W"], tmp_path)
        assert hashlib.md5(out).hexdigest() == case["out_md5"]
        assert so.decode() == case["stdout"] and se.decode() == case["stderr"]


def test_msa_driver_reference_fuzz_goldens(gpu, tmp_path):
    g = json.load(open(GOLDEN / "msa_golden.json"))
    for case in g["fuzz"]:
        out, _, _ = run_msa([r.encode() for r in case["reads"]], case["params"], case["W"], tmp_path)
        assert hashlib.md5(out).hexdigest() == case["out_md5"], case["name"]
