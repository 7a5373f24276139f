"""Parity of the CUDA pairwise path (through the C ABI) with the oracle and the reference's
goldens.  Bit-exact: maxsorce, corner score, last DP row, every 4-bit direction code and the
bytes of the -o file."""
import hashlib
import json
import random

import numpy as np
import pytest

from oracle import binding as ob
from tsta_b200 import synth
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

PARAMS = [(2, -5, -2, -4), (2, -3, -2, -4), (1, -1, -1, 0), (2, -5, -2, 0), (5, -4, -1, -10),
          (1, -3, -1, -5), (2, -5, 0, 0), (10, -10, -5, -20), (60, -5, -2, -4), (2, -60, -2, -4)]


def rescore(a: bytes, b: bytes, M, X, E, O) -> int:
    """Affine score of an emitted alignment (gap of length k costs O + k*E)."""
    A = np.frombuffer(a, dtype=np.uint8)
    B = np.frombuffer(b, dtype=np.uint8)
    ga, gb = A == ord("-"), B == ord("-")
    both = ~ga & ~gb
    s = int(np.sum(np.where(A[both] == B[both], M, X)))
    for g in (ga, gb):
        s += int(g.sum()) * E
        opens = int(np.sum(g[1:] & ~g[:-1])) + int(g[0])
        s += opens * O
    return s


def test_example_pair_golden(gpu):
    g = json.load(open(GOLDEN / "psa_example.json"))
    a, b = g["seqa"].encode(), g["seqb"].encode()
    for case in g["cases"]:
        M, X, E, O = case["params"]
        r = gpu.psa(a, b, M, X, E, O, trace=True)
        assert r.maxsorce == case["maxsorce"]
        assert r.corner == case["corner"]
        assert hashlib.md5(r.output_bytes()).hexdigest() == case["out_md5"]
        r0 = gpu.psa(a, b, M, X, E, O, trace=False)
        assert (r0.maxsorce, r0.corner) == (case["maxsorce"], case["corner"])


def test_committed_reference_goldens(gpu):
    for c in json.load(open(GOLDEN / "psa_golden.json")):
        x, y = c["x"].encode(), c["y"].encode()
        M, X, E, O = c["params"]
        r = gpu.psa(x, y, M, X, E, O, trace=True)
        assert r.maxsorce == c["maxsorce"], c["name"]
        assert hashlib.md5(r.output_bytes()).hexdigest() == c["out_md5"], c["name"]
        assert gpu.psa(x, y, M, X, E, O, trace=False).maxsorce == c["maxsorce"], c["name"]


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_full_state_vs_oracle(gpu, seed):
    """Every intermediate the ABI can export: last row, all direction codes, strings."""
    rnd = random.Random(seed)
    lens = [1, 2, 31, 32, 33, 127, 128, 129, 255, 256, 257, 500, 1023, 1025]
    plan_t, plan_s = gpu.PsaPlan(True), gpu.PsaPlan(False)
    for k in range(10):
        la, lb = rnd.choice(lens), rnd.choice(lens)
        x, y = synth.read_pair(la, lb, 7000 + 100 * seed + k, rnd.choice([None, 0.15, 0.03]))
        y = y or b"T"
        M, X, E, O = rnd.choice(PARAMS)
        o = ob.psa(x, y, M, X, E, O, want_lastrow=True, want_dirs=True)
        for plan in (plan_t, plan_s):
            plan.upload(x, y, M, X, E, O)
            plan.run()
            r = plan.fetch()
            assert (r.maxsorce, r.corner) == (o.maxsorce, o.corner), (la, lb, M, X, E, O)
            np.testing.assert_array_equal(plan.lastrow(), o.lastrow)
        np.testing.assert_array_equal(plan_t.dirs(), o.dirs)
        r = plan_t.fetch()
        assert r.output_bytes() == o.output_bytes()
    plan_t.close()
    plan_s.close()


def test_mid_size_pairs(gpu):
    for k, (la, lb, div) in enumerate([(5000, 4000, None), (6001, 6001, 0.05), (12000, 700, None),
                                       (20000, 19000, 0.1)]):
        x, y = synth.read_pair(la, lb, 900 + k, div)
        for prm in [(2, -5, -2, -4), (2, -5, -2, 0)]:
            o = ob.psa(x, y, *prm)
            r = gpu.psa(x, y, *prm, trace=True)
            assert (r.maxsorce, r.corner) == (o.maxsorce, o.corner)
            assert r.output_bytes() == o.output_bytes()
            r0 = gpu.psa(x, y, *prm, trace=False)
            assert (r0.maxsorce, r0.corner) == (o.maxsorce, o.corner)


def test_many_symbols_uses_generic_compare_path(gpu):
    rnd = random.Random(3)
    alpha = b"ACDEFGHIKLMNPQRSTVWY"
    x = bytes(rnd.choice(alpha) for _ in range(700))
    y = bytes(rnd.choice(alpha) for _ in range(650))
    for prm in [(5, -4, -1, -10), (2, -5, -2, -4)]:
        o = ob.psa(x, y, *prm)
        r = gpu.psa(x, y, *prm)
        assert (r.maxsorce, r.corner) == (o.maxsorce, o.corner)
        assert r.output_bytes() == o.output_bytes()
    # lower/upper case are different symbols (raw byte equality, psa.c:253)
    o = ob.psa(b"ACGTacgtNNacgtACGT" * 20, b"acgtACGTnnACGTacgt" * 19)
    r = gpu.psa(b"ACGTacgtNNacgtACGT" * 20, b"acgtACGTnnACGTacgt" * 19)
    assert r.output_bytes() == o.output_bytes() and r.maxsorce == o.maxsorce


@pytest.mark.parametrize("n,cols,packed", [(120_000, 8, False), (303_200, 16, False), (120_000, 4, True), (303_200, 8, True)])
def test_wide_strip_variants_long_s0(gpu, n, cols, packed, monkeypatch):
    """Longer s0 switches lanes to wider strips (int32: 8 / 16 columns, packed: 4 / 8); few
    rows keep the oracle fast."""
    monkeypatch.setenv("TSTA_PSA_PACKED", "1" if packed else "0")
    x, y = synth.read_pair(n, 40, 42)
    o = ob.psa(x, y, want_lastrow=True)
    plan = gpu.PsaPlan(True)
    plan.upload(x, y)
    plan.run()
    assert plan.info()["cols_per_lane"] == cols and plan.info()["packed"] == packed
    r = plan.fetch()
    assert (r.maxsorce, r.corner) == (o.maxsorce, o.corner)
    np.testing.assert_array_equal(plan.lastrow(), o.lastrow)
    assert r.output_bytes() == o.output_bytes()
    plan.close()


def test_argument_errors(gpu):
    with pytest.raises(gpu.TstaError):
        gpu.psa(b"ACGT", b"ACG", O=3)          # O > 0 is outside the parity domain
    import ctypes as C
    L = gpu.load()
    res = gpu._PsaResult()
    p = gpu.Params(2, -5, -2, -4)
    assert L.tsta_psa_align(C.byref(p), b"AC", 2, b"ACGT", 4, 1, C.byref(res)) == -1  # n < m
    assert b"longer" in L.tsta_last_error()


def test_full_size_properties_100k(gpu):
    """BASELINE config 2/3 scale (1e10 cells): size-independent properties."""
    n = 100_000
    x = synth.random_read(n, 1)
    # (1) self alignment: all diagonal, score M*n
    r = gpu.psa(x, x, trace=True)
    assert r.corner == 2 * n and r.maxsorce == 2 * n
    assert r.a == x and r.b == x
    # (2) 5%-divergent copy: the emitted alignment spells both reads and rescoring it gives
    #     exactly the corner score the fill reported; score-only and traced fills agree
    y = synth.mutate(x, 0.05, 9)
    s0, s1 = ob.order_pair(x, y)
    r = gpu.psa(x, y, trace=True)
    assert r.a.replace(b"-", b"") == s0 and r.b.replace(b"-", b"") == s1
    assert rescore(r.a, r.b, 2, -5, -2, -4) == r.corner
    r0 = gpu.psa(x, y, trace=False)
    assert (r0.maxsorce, r0.corner) == (r.maxsorce, r.corner)
    # (3) linear gaps
    r = gpu.psa(x, y, O=0, trace=True)
    assert rescore(r.a, r.b, 2, -5, -2, 0) == r.corner


def test_full_size_properties_200k_traced(gpu):
    """BASELINE config 3 (4e10 cells, 20 GB of packed directions): the emitted alignment spells
    both reads, rescoring it gives the corner score the fill reported, the score-only fill agrees,
    and the walk of a related pair stays exact over 200k rows."""
    n = 200_000
    x = synth.random_read(n, 21)
    y = synth.mutate(x, 0.08, 5)[:n]
    s0, s1 = ob.order_pair(x, y)
    r = gpu.psa(x, y, trace=True)
    assert r.a.replace(b"-", b"") == s0 and r.b.replace(b"-", b"") == s1
    assert rescore(r.a, r.b, 2, -5, -2, -4) == r.corner
    r0 = gpu.psa(x, y, trace=False)
    assert (r0.maxsorce, r0.corner) == (r.maxsorce, r.corner)
    # unrelated pair of the benchmark shape: alignment is valid and optimal w.r.t. the reported corner
    u, v = synth.random_read(n, 1), synth.random_read(n, 2)
    r = gpu.psa(u, v, trace=True)
    assert r.a.replace(b"-", b"") == u and r.b.replace(b"-", b"") == v
    assert rescore(r.a, r.b, 2, -5, -2, -4) == r.corner
