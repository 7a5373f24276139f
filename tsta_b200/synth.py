"""Deterministic synthetic reads for parity tests and bench.py.

xorshift64* streams, `"ACGT"[x & 3]`; mutated copies use equal shares of substitution,
insertion and deletion.  numpy only -- no dependence on the CUDA library or the oracle.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _xorshift64star(seed: int, count: int) -> np.ndarray:
    """Vectorised-by-blocks xorshift64*: `count` uint64 outputs of the stream `seed`."""
    # The scalar recurrence is inherently serial; run 1024 independent lanes seeded from
    # a splitmix of the seed so large reads (1e5..1e6 bases) generate in milliseconds.
    lanes = 1024
    s = np.arange(1, lanes + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)
    s ^= s >> np.uint64(30)
    s *= np.uint64(0xBF58476D1CE4E5B9)
    s ^= s >> np.uint64(27)
    s *= np.uint64(0x94D049BB133111EB)
    s ^= s >> np.uint64(31)
    s[s == 0] = np.uint64(0x2545F4914F6CDD1D)
    rounds = (count + lanes - 1) // lanes
    out = np.empty((rounds, lanes), dtype=np.uint64)
    for r in range(rounds):
        s ^= s >> np.uint64(12)
        s ^= s << np.uint64(25)
        s ^= s >> np.uint64(27)
        out[r] = s * np.uint64(0x2545F4914F6CDD1D)
    return out.reshape(-1)[:count]


def random_read(length: int, seed: int) -> bytes:
    """Uniform random ACGT read of exactly `length` bases."""
    if length <= 0:
        return b""
    x = _xorshift64star(seed, length)
    return _ACGT[((x >> np.uint64(33)) & np.uint64(3)).astype(np.intp)].tobytes()


def mutate(read: bytes, rate: float, seed: int) -> bytes:
    """Copy of `read` with ~`rate` edits per base: 1/3 substitution, 1/3 insertion, 1/3 deletion."""
    n = len(read)
    if n == 0 or rate <= 0:
        return read
    x = _xorshift64star(seed ^ 0x5bd1e995, 2 * n)
    u = (x[:n] >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    kind = ((x[n:] >> np.uint64(20)) % np.uint64(3)).astype(np.int64)
    base = _ACGT[((x[n:] >> np.uint64(40)) & np.uint64(3)).astype(np.intp)]
    src = np.frombuffer(read, dtype=np.uint8)
    hit = u < rate
    out = bytearray()
    last = 0
    for p in np.nonzero(hit)[0]:
        out += src[last:p].tobytes()
        k = kind[p]
        if k == 0:      # substitution (forced to differ)
            c = base[p]
            if c == src[p]:
                c = _ACGT[(int(np.searchsorted(_ACGT, c)) + 1) & 3] if c in _ACGT else _ACGT[0]
            out.append(int(c))
        elif k == 1:    # insertion before p
            out.append(int(base[p]))
            out.append(int(src[p]))
        # k == 2: deletion -> emit nothing
        last = p + 1
    out += src[last:].tobytes()
    return bytes(out)


def read_pair(length_a: int, length_b: int, seed: int, divergence: float | None = None):
    """(a, b): independent random reads, or b = mutated copy of a when divergence is given."""
    a = random_read(length_a, 2 * seed + 1)
    if divergence is None:
        b = random_read(length_b, 2 * seed + 2)
    else:
        b = mutate(a, divergence, seed)
    return a, b


def random_bulk(count: int, seed: int) -> np.ndarray:
    """`count` uniform random ACGT bases as a uint8 array, for the batch workloads (1e8..1e9 bases):
    numpy's PCG64 stream instead of the xorshift64* lanes above, which generate ~20 MB/s."""
    if count <= 0:
        return np.zeros(0, dtype=np.uint8)
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty(count, dtype=np.uint8)
    step = 1 << 26
    for o in range(0, count, step):
        k = min(step, count - o)
        out[o:o + k] = _ACGT[rng.integers(0, 4, size=k, dtype=np.uint8)]
    return out
