"""Host-side helpers that need no GPU: column slicing for one pair over several GPUs and the
assembly of the -o strings from walker moves (tsta_b200/api.py), against the oracle's output."""
import random

import numpy as np

from oracle import binding as ob
from tsta_b200 import api, synth


def moves_from_alignment(a: bytes, b: bytes) -> np.ndarray:
    """Walker moves (end of the alignment first): 1 diagonal, 0 gap in s1, 2 gap in s0."""
    A = np.frombuffer(a, dtype=np.uint8)
    B = np.frombuffer(b, dtype=np.uint8)
    ops = np.where(B == ord("-"), 0, np.where(A == ord("-"), 2, 1)).astype(np.uint8)
    return ops[::-1]


def test_assemble_alignment_reproduces_oracle_strings():
    rnd = random.Random(2)
    for k in range(12):
        x, y = synth.read_pair(rnd.choice([1, 5, 40, 300]), rnd.choice([1, 7, 33, 280]), 60 + k, rnd.choice([None, 0.1]))
        s0, s1 = ob.order_pair(x, y or b"C")
        o = ob.psa(x, y or b"C")
        out = o.output_bytes()
        a, b = out[len(b">1\n"):].split(b"\n>2\n")
        mv = moves_from_alignment(a, b)
        assert api.assemble_alignment(s0, s1, mv) == (a, b)
        # the walker stops at the matrix edge: the trailing run (leading columns of the alignment) of
        # "rest of s0 against gaps, then rest of s1" is appended by assemble_alignment itself (psa.c:497-510)
        cut = len(mv)
        while cut > 0 and mv[cut - 1] == 2:
            cut -= 1
        while cut > 0 and mv[cut - 1] == 0:
            cut -= 1
        assert api.assemble_alignment(s0, s1, mv[:cut]) == (a, b)


def test_slice_columns_cover_and_granule():
    for n in [1, 1535, 1536, 1537, 100_000, 400_000, 1_000_003]:
        for parts in [1, 2, 3, 4, 8]:
            sl = api.slice_columns(n, parts)
            assert 1 <= len(sl) <= parts
            assert sl[0][0] == 0 and sum(w for _, w in sl) == n
            for (c0, w), (c1, _) in zip(sl, sl[1:]):
                assert c0 + w == c1 and w % 1536 == 0      # every slice with a right neighbour: whole bands
    assert api.slice_columns(0, 4) == [] and api.slice_columns(10, 0) == []
