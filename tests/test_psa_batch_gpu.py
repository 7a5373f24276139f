"""Batched pairwise path (tsta_psa_batch): every pair must equal what the single-pair path /
the oracle gives for that pair alone."""
import random

import numpy as np
import pytest

from oracle import binding as ob
from tsta_b200 import synth

pytestmark = pytest.mark.gpu


def _mixed_pairs(count, seed, lens):
    rnd = random.Random(seed)
    pairs = []
    for k in range(count):
        la, lb = rnd.choice(lens), rnd.choice(lens)
        x, y = synth.read_pair(la, lb, seed * 1000 + k, rnd.choice([None, 0.05, 0.15]))
        pairs.append((x, y or b"A"))
    return pairs


@pytest.mark.parametrize("prm", [(2, -5, -2, -4), (2, -5, -2, 0), (5, -4, -1, -10)])
def test_short_read_batch_vs_oracle(gpu, prm):
    pairs = _mixed_pairs(300, 3, [1, 2, 30, 100, 127, 128, 129, 200, 250, 255, 256, 257, 300])
    r0 = gpu.psa_batch(pairs, *prm, trace=False)
    r1 = gpu.psa_batch(pairs, *prm, trace=True)
    for k, (x, y) in enumerate(pairs):
        o = ob.psa(x, y, *prm)
        assert (r0.maxsorce[k], r0.corner[k]) == (o.maxsorce, o.corner), k
        assert (r1.maxsorce[k], r1.corner[k]) == (o.maxsorce, o.corner), k
        assert r1.output_bytes(k) == o.output_bytes(), k


@pytest.mark.parametrize("chunks", [3, 8])
def test_chunked_upload_mixed_alphabets(gpu, chunks, monkeypatch):
    """Large batches are uploaded in chunks that overlap with the kernels of the previous chunk; every
    chunk scans its own alphabet, so a batch whose chunks differ (ACGT reads, then 20-letter reads)
    runs packed kernels on one chunk and the generic ones on the other."""
    monkeypatch.setenv("TSTA_BATCH_CHUNKS", str(chunks))
    rnd = random.Random(4)
    alpha = b"ACDEFGHIKLMNPQRSTVWY"
    pairs = _mixed_pairs(150, 8, [1, 60, 128, 250, 256])
    pairs += [(bytes(rnd.choice(alpha) for _ in range(rnd.randint(1, 200))),
               bytes(rnd.choice(alpha) for _ in range(rnd.randint(1, 200)))) for _ in range(90)]
    pairs += _mixed_pairs(60, 9, [100, 250])
    r0 = gpu.psa_batch(pairs, trace=False)
    r1 = gpu.psa_batch(pairs, trace=True)
    for k, (x, y) in enumerate(pairs):
        o = ob.psa(x, y)
        assert (r0.maxsorce[k], r0.corner[k]) == (o.maxsorce, o.corner), k
        assert (r1.maxsorce[k], r1.corner[k]) == (o.maxsorce, o.corner), k
        assert r1.output_bytes(k) == o.output_bytes(), k


def test_long_pairs_in_a_batch(gpu):
    pairs = _mixed_pairs(12, 5, [3000, 5000, 10000]) + _mixed_pairs(40, 6, [250])
    r1 = gpu.psa_batch(pairs, trace=True)
    r0 = gpu.psa_batch(pairs, trace=False)
    for k, (x, y) in enumerate(pairs):
        o = ob.psa(x, y)
        assert (r0.maxsorce[k], r0.corner[k]) == (o.maxsorce, o.corner), k
        assert (r1.maxsorce[k], r1.corner[k]) == (o.maxsorce, o.corner), k
        assert r1.output_bytes(k) == o.output_bytes(), k


def test_narrow_strip_batch_and_many_symbols(gpu):
    rnd = random.Random(9)
    alpha = b"ACDEFGHIKLMNPQRSTVWY"
    pairs = [(bytes(rnd.choice(alpha) for _ in range(rnd.randint(1, 128))),
              bytes(rnd.choice(alpha) for _ in range(rnd.randint(1, 128)))) for _ in range(100)]
    r = gpu.psa_batch(pairs, 5, -4, -1, -10, trace=True)
    for k, (x, y) in enumerate(pairs):
        o = ob.psa(x, y, 5, -4, -1, -10)
        assert (r.maxsorce[k], r.corner[k]) == (o.maxsorce, o.corner), k
        assert r.output_bytes(k) == o.output_bytes(), k


def test_batch_equals_single_pair_path_at_scale(gpu):
    """20k pairs of 250 bp: spot-check against the oracle, checksum-compare score-only vs traced."""
    n = 20000
    a = synth.random_read(250 * n, 77)
    pairs = []
    for k in range(n):
        x = a[250 * k: 250 * (k + 1)]
        pairs.append((x, synth.mutate(x, 0.05, k) if k % 2 else synth.random_read(250, 10_000 + k)))
    r0 = gpu.psa_batch(pairs, trace=False)
    r1 = gpu.psa_batch(pairs, trace=True)
    np.testing.assert_array_equal(r0.maxsorce, r1.maxsorce)
    np.testing.assert_array_equal(r0.corner, r1.corner)
    for k in range(0, n, 397):
        o = ob.psa(*pairs[k])
        assert (r0.maxsorce[k], r0.corner[k]) == (o.maxsorce, o.corner)
        assert r1.output_bytes(k) == o.output_bytes()


@pytest.mark.parametrize("chunks", [1, 4])
def test_packed_moves_equal_byte_moves(gpu, chunks, monkeypatch):
    """tsta_psa_batch_packed (four moves per byte, what travels back over PCIe in bench.py) decodes to the
    alignments of the byte form and of the oracle, for ragged lengths and chunked downloads."""
    monkeypatch.setenv("TSTA_BATCH_CHUNKS", str(chunks))
    pairs = _mixed_pairs(400, 11, [1, 2, 3, 5, 30, 100, 127, 128, 129, 250, 255, 256, 257, 300, 700])
    rb = gpu.psa_batch(pairs, trace=True)
    rp = gpu.psa_batch(pairs, trace=True, packed=True)
    assert (rb.maxsorce == rp.maxsorce).all() and (rb.corner == rp.corner).all() and (rb._alen == rp._alen).all()
    assert rp._ops.nbytes * 3 < rb._ops.nbytes
    for k, (x, y) in enumerate(pairs):
        assert rp.output_bytes(k) == rb.output_bytes(k), k
        if k % 7 == 0:
            assert rp.output_bytes(k) == ob.psa(x, y).output_bytes(), k


@pytest.mark.parametrize("chunks", [1, 3])
def test_2bit_wire_format_equals_byte_reads(gpu, chunks, monkeypatch):
    """tsta_psa_batch_2bit (reads four bases to a byte over PCIe, expanded on the device; chunk boundaries fall inside
    bytes) against the byte form and the oracle, score-only and traced, ragged lengths."""
    monkeypatch.setenv("TSTA_BATCH_CHUNKS", str(chunks))
    pairs = _mixed_pairs(500, 23, [1, 2, 3, 5, 30, 99, 127, 128, 129, 250, 255, 256, 257, 301, 700])
    for trace in (False, True):
        rb = gpu.psa_batch(pairs, trace=trace, packed=trace)
        r2 = gpu.psa_batch(pairs, trace=trace, bits2=True)
        assert (rb.maxsorce == r2.maxsorce).all() and (rb.corner == r2.corner).all()
        for k, (x, y) in enumerate(pairs):
            if trace:
                assert r2.output_bytes(k) == rb.output_bytes(k), k
            if k % 9 == 0:
                o = ob.psa(x, y)
                assert (r2.maxsorce[k], r2.corner[k]) == (o.maxsorce, o.corner), k
                if trace:
                    assert r2.output_bytes(k) == o.output_bytes(), k


def test_bad_descriptors_are_rejected_with_the_first_bad_pair(gpu):
    """The descriptor scan runs on the device (batch_scan_kernel): a batch with a bad pair fails with TSTA_EINVAL naming
    the FIRST bad pair, nothing is computed; the call after it works (psa/psa.c:194-199: the caller orders the pair)."""
    n, ln = 5000, 40
    rng = np.random.default_rng(5)
    s0 = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), n * ln)
    s1 = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), n * ln)
    off = np.arange(n, dtype=np.int64) * ln
    l0 = np.full(n, ln, dtype=np.int32)
    l1 = np.full(n, ln, dtype=np.int32)

    def run(o0=off, a=l0, o1=off, c=l1, trace=False, packed=False, out=None):
        return gpu.psa_batch_arrays(s0, o0, a, s1, o1, c, trace=trace, packed=packed, out=out)

    good = run()
    short0 = l0.copy(); short0[[4321, 777]] = ln - 1              # len0 < len1 at two pairs: the lower index is reported
    with pytest.raises(gpu.TstaError, match="pair 777:"):
        run(a=short0)
    neg = off.copy(); neg[3000] = -8
    with pytest.raises(gpu.TstaError, match="pair 3000:"):
        run(o1=neg)
    neg_len = l1.copy(); neg_len[12] = -1
    with pytest.raises(gpu.TstaError, match="pair 12:"):
        run(c=neg_len)
    # packed moves: every slot has to start at a multiple of four moves
    cap = 2 * ln
    ops_off = np.arange(n, dtype=np.int64) * cap
    bad_off = ops_off.copy(); bad_off[99] += 2
    out = (np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n * cap // 4 + 4, np.uint8), bad_off, np.zeros(n, np.int32))
    with pytest.raises(gpu.TstaError, match="pair 99:"):
        run(trace=True, packed=True, out=out)
    again = run()
    assert np.array_equal(again.maxsorce, good.maxsorce) and np.array_equal(again.corner, good.corner)
    k = 1234
    o = ob.psa(s0[k * ln:(k + 1) * ln].tobytes(), s1[k * ln:(k + 1) * ln].tobytes())
    assert (good.maxsorce[k], good.corner[k]) == (o.maxsorce, o.corner)


def test_release_cache_returns_the_batch_context(gpu):
    """tsta_psa_release_cache() also frees what tsta_psa_batch keeps between calls; the next call rebuilds it."""
    pairs = _mixed_pairs(60, 9, [30, 100, 250])
    r0 = gpu.psa_batch(pairs, trace=True)
    gpu.load().tsta_psa_release_cache()
    gpu.load().tsta_psa_release_cache()          # (idempotent)
    r1 = gpu.psa_batch(pairs, trace=True)
    assert np.array_equal(r0.maxsorce, r1.maxsorce) and np.array_equal(r0.corner, r1.corner)
    assert all(r0.output_bytes(k) == r1.output_bytes(k) for k in range(len(pairs)))
