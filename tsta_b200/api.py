"""Python mirror of the C ABI in include/tsta_cuda.h (ctypes; no torch types cross it).

The reference has no library API -- its interface is the TSTA_psa / TSTA_psa_notrace /
TSTA_msa executables -- so this module mirrors those at function level for tests and
bench.py: `psa(x, y, ...)` does what `TSTA_psa -1 x.fa -2 y.fa` does (swap so the longer
read is first, psa/psa.c:194-199; fill; traceback) and returns the printed `maxsorce` plus
the bytes of the `-o` file.

There is no CPU fallback: if libtsta_cuda.so is missing or no CUDA device is usable every
call raises TstaError.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
import os as _os
LIB_PATH = Path(_os.environ["TSTA_LIB"]) if _os.environ.get("TSTA_LIB") else _PKG / "_build" / "libtsta_cuda.so"   # (TSTA_LIB: experiment builds)
_lib = None


class TstaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"tsta error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("M", C.c_int32), ("X", C.c_int32), ("E", C.c_int32), ("O", C.c_int32)]


class _PsaResult(C.Structure):
    _fields_ = [("maxsorce", C.c_int32), ("corner", C.c_int32), ("alen", C.c_int64),
                ("a", C.c_void_p), ("b", C.c_void_p)]


def load() -> C.CDLL:
    """dlopen the in-tree CUDA library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise TstaError(-2, f"{LIB_PATH} not built: run `python -m tsta_b200.build` (needs nvcc); "
                            "there is no CPU fallback")
    L = C.CDLL(str(LIB_PATH))
    L.tsta_last_error.restype = C.c_char_p
    L.tsta_version.restype = C.c_char_p
    L.tsta_psa_align.argtypes = [C.POINTER(Params), C.c_char_p, C.c_int64, C.c_char_p, C.c_int64,
                                 C.c_int, C.POINTER(_PsaResult)]
    L.tsta_psa_result_free.argtypes = [C.POINTER(_PsaResult)]
    L.tsta_psa_plan_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.tsta_psa_plan_upload.argtypes = [C.c_void_p, C.POINTER(Params), C.c_char_p, C.c_int64,
                                       C.c_char_p, C.c_int64]
    L.tsta_psa_plan_run.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.tsta_psa_plan_fetch.argtypes = [C.c_void_p, C.POINTER(_PsaResult)]
    L.tsta_psa_plan_lastrow.argtypes = [C.c_void_p, C.c_void_p]
    L.tsta_psa_plan_dirs.argtypes = [C.c_void_p, C.c_void_p]
    L.tsta_psa_plan_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.tsta_psa_plan_destroy.argtypes = [C.c_void_p]
    L.tsta_psa_plan_packed.argtypes = [C.c_void_p]
    L.tsta_psa_plan_launch.argtypes = [C.c_void_p]
    L.tsta_psa_plan_wait.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.tsta_psa_plan_upload_slice.argtypes = [C.c_void_p, C.POINTER(Params), C.c_char_p, C.c_int64, C.c_int64, C.c_int64,
                                             C.c_char_p, C.c_int64, C.c_int, C.c_int]
    L.tsta_psa_plan_inbox_handle.argtypes = [C.c_void_p, C.c_char_p]
    L.tsta_psa_plan_connect_right.argtypes = [C.c_void_p, C.c_char_p]
    L.tsta_psa_plan_connect_right_local.argtypes = [C.c_void_p, C.c_void_p]
    L.tsta_psa_plan_trace_from.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int),
                                           C.POINTER(C.c_int64), C.POINTER(C.c_float)]
    L.tsta_psa_plan_ops.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.tsta_set_device.argtypes = [C.c_int]
    L.tsta_measure_dpx_peak.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_float)]
    if hasattr(L, "tsta_psa_batch"):
        L.tsta_psa_batch.argtypes = [
            C.POINTER(Params), C.c_int64,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.POINTER(C.c_float)]
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise TstaError(rc, load().tsta_last_error().decode())


def device_count() -> int:
    n = load().tsta_device_count()
    if n < 0:
        raise TstaError(n, load().tsta_last_error().decode())
    return n


def set_device(dev: int) -> None:
    _check(load().tsta_set_device(dev))


def order_pair(x: bytes, y: bytes):
    """psa/psa.c:194-199: the longer read becomes s0 (columns); equal lengths keep order."""
    return (y, x) if len(y) > len(x) else (x, y)


@dataclass
class PsaResult:
    maxsorce: int
    corner: int
    a: bytes | None = None
    b: bytes | None = None

    def output_bytes(self) -> bytes:
        """Bytes of the TSTA_psa -o file (psa/psa.c:512-515: no trailing newline)."""
        return b">1\n" + self.a + b"\n>2\n" + self.b


def _take(res: _PsaResult) -> PsaResult:
    out = PsaResult(res.maxsorce, res.corner)
    if res.a:
        out.a = C.string_at(res.a, res.alen)
        out.b = C.string_at(res.b, res.alen)
    load().tsta_psa_result_free(C.byref(res))
    return out


def psa(x: bytes, y: bytes, M=2, X=-5, E=-2, O=-4, trace=True) -> PsaResult:
    """One pair through the one-shot C entry point (host buffers in, host strings out)."""
    s0, s1 = order_pair(x, y)
    p = Params(M, X, E, O)
    res = _PsaResult()
    _check(load().tsta_psa_align(C.byref(p), s0, len(s0), s1, len(s1), int(trace), C.byref(res)))
    return _take(res)


class PsaPlan:
    """Device-resident pair: upload once, run many times (what bench.py times)."""

    def __init__(self, trace: bool):
        self._h = C.c_void_p()
        self.trace = bool(trace)
        _check(load().tsta_psa_plan_create(C.byref(self._h), int(trace)))
        self.n = self.m = 0

    def upload(self, x: bytes, y: bytes, M=2, X=-5, E=-2, O=-4) -> None:
        s0, s1 = order_pair(x, y)
        self.n, self.m = len(s0), len(s1)
        p = Params(M, X, E, O)
        _check(load().tsta_psa_plan_upload(self._h, C.byref(p), s0, self.n, s1, self.m))

    def upload_slice(self, s0: bytes, s1: bytes, col0: int, ncols: int, M=2, X=-5, E=-2, O=-4) -> None:
        """Columns [col0, col0+ncols) of the pair (s0 longer, already ordered) on this plan's GPU."""
        self.n, self.m = ncols, len(s1)
        p = Params(M, X, E, O)
        _check(load().tsta_psa_plan_upload_slice(self._h, C.byref(p), s0[col0:col0 + ncols], ncols, col0, len(s0),
                                                 s1, len(s1), int(col0 > 0), int(col0 + ncols < len(s0))))

    def inbox_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        _check(load().tsta_psa_plan_inbox_handle(self._h, buf))
        return buf.raw

    def connect_right(self, handle: bytes) -> None:
        _check(load().tsta_psa_plan_connect_right(self._h, handle))

    def connect_right_local(self, right: "PsaPlan") -> None:
        _check(load().tsta_psa_plan_connect_right_local(self._h, right._h))

    def trace_from(self, start_i: int, start_forced: int = 0):
        """Walk this (slice's) direction matrix; returns (moves uint8[nops], end_i, end_forced, ms)."""
        ei, ef, n, ms = C.c_int64(), C.c_int(), C.c_int64(), C.c_float()
        _check(load().tsta_psa_plan_trace_from(self._h, start_i, start_forced, C.byref(ei), C.byref(ef), C.byref(n), C.byref(ms)))
        ops = np.zeros(max(n.value, 1), dtype=np.uint8)
        _check(load().tsta_psa_plan_ops(self._h, ops.ctypes.data, n.value))
        return ops[:n.value], ei.value, ef.value, ms.value

    def launch(self) -> None:
        _check(load().tsta_psa_plan_launch(self._h))

    def wait(self):
        f, t = C.c_float(), C.c_float()
        _check(load().tsta_psa_plan_wait(self._h, C.byref(f), C.byref(t)))
        return f.value, t.value

    def run(self):
        """Launch fill (+ traceback); returns (fill_ms, trace_ms) from CUDA events."""
        f, t = C.c_float(), C.c_float()
        _check(load().tsta_psa_plan_run(self._h, C.byref(f), C.byref(t)))
        return f.value, t.value

    def fetch(self) -> PsaResult:
        res = _PsaResult()
        _check(load().tsta_psa_plan_fetch(self._h, C.byref(res)))
        return _take(res)

    def lastrow(self) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.int32)
        _check(load().tsta_psa_plan_lastrow(self._h, out.ctypes.data))
        return out

    def dirs(self) -> np.ndarray:
        out = np.zeros((self.m, self.n), dtype=np.uint8)
        _check(load().tsta_psa_plan_dirs(self._h, out.ctypes.data))
        return out

    def info(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        _check(load().tsta_psa_plan_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"launches": a.value, "cols_per_lane": b.value, "bands": c.value,
                "packed": bool(load().tsta_psa_plan_packed(self._h))}

    def trace_stats(self):
        """Checkpointed traceback of the last run: corridors used, tiles re-run, tiles of the whole matrix."""
        a, b, c = C.c_int(), C.c_longlong(), C.c_longlong()
        L = load()
        L.tsta_psa_plan_trace_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _check(L.tsta_psa_plan_trace_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"corridors": a.value, "tiles": b.value, "tiles_of_matrix": c.value}

    def close(self) -> None:
        if self._h:
            load().tsta_psa_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def assemble_alignment(s0: bytes, s1: bytes, moves: np.ndarray):
    """(a, b) strings of TSTA_psa's -o file from the moves in walk order (end of the alignment first),
    including the leftovers of psa/psa.c:497-510 (rest of s0 against gaps, then rest of s1)."""
    n0 = int(np.count_nonzero(moves != 2))      # bases of s0 consumed
    n1 = int(np.count_nonzero(moves != 0))
    tail = np.concatenate([moves, np.zeros(len(s0) - n0, dtype=np.uint8), np.full(len(s1) - n1, 2, dtype=np.uint8)])
    ops = tail[::-1]
    A = np.full(len(ops), ord("-"), dtype=np.uint8)
    B = np.full(len(ops), ord("-"), dtype=np.uint8)
    A[ops != 2] = np.frombuffer(s0, dtype=np.uint8)
    B[ops != 0] = np.frombuffer(s1, dtype=np.uint8)
    return A.tobytes(), B.tobytes()


def slice_columns(n: int, parts: int, granule: int = 1536):
    """Column slices of one long pair for `parts` GPUs: every slice but the last is a multiple of
    `granule` columns (a whole number of bands for every strip width).  Returns [(col0, ncols)];
    fewer than `parts` entries when the read is too short to give every GPU a slice."""
    if n <= 0 or parts <= 0:
        return []
    width = -(-(-(-n // parts)) // granule) * granule
    out, c = [], 0
    while c < n:
        w = min(width, n - c)
        out.append((c, w))
        c += w
    return out


class PinnedArray:
    """numpy view of page-locked host memory (tsta_host_alloc); keep the object alive while in use."""

    def __init__(self, shape, dtype):
        L = load()
        L.tsta_host_alloc.restype = C.c_void_p
        L.tsta_host_alloc.argtypes = [C.c_size_t]
        L.tsta_host_free.argtypes = [C.c_void_p]
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
        self._p = L.tsta_host_alloc(max(n * dt.itemsize, 1))
        if not self._p:
            raise TstaError(-3, L.tsta_last_error().decode())
        buf = (C.c_char * max(n * dt.itemsize, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dt, count=n).reshape(shape)

    def __del__(self):
        try:
            if self._p:
                load().tsta_host_free(self._p)
                self._p = None
        except Exception:
            pass


def pinned_copy(a: np.ndarray) -> PinnedArray:
    p = PinnedArray(a.shape, a.dtype)
    p.array[...] = a
    return p


def measure_dpx_peak():
    """(int32 DPX lane-ops per second, kernel ms) from the on-device microbenchmark."""
    v, ms = C.c_double(), C.c_float()
    _check(load().tsta_measure_dpx_peak(C.byref(v), C.byref(ms)))
    return v.value, ms.value


@dataclass
class BatchResult:
    maxsorce: np.ndarray            # int32 [npairs]
    corner: np.ndarray              # int32 [npairs]
    kernel_ms: float
    swapped: np.ndarray             # bool [npairs]: pair k was swapped so that s0 is the longer read
    _ops: np.ndarray | None = None
    _ops_off: np.ndarray | None = None
    _alen: np.ndarray | None = None
    _pairs: list | None = None
    _packed: bool = False           # moves packed four to a byte (tsta_psa_batch_packed)

    def moves(self, k: int, cap: int) -> np.ndarray:
        """The alen[k] moves of pair k in reading order (1 both bases, 0 s0 base against a gap, 2 gap against an s1 base);
        cap = len0 + len1 of the pair (the size of its slot)."""
        n = int(self._alen[k])
        o = int(self._ops_off[k])
        if not self._packed:
            return self._ops[o + cap - n: o + cap]
        b0 = (o + cap - n) >> 2
        raw = self._ops[b0: ((o + cap + 3) >> 2)]
        two = ((raw[:, None] >> np.array([0, 2, 4, 6], dtype=np.uint8)) & 3).reshape(-1)
        first = (o + cap - n) - 4 * b0
        return two[first: first + n]

    def alignment(self, k: int):
        """(a, b) strings of pair k exactly as TSTA_psa would write them (">1" = longer read)."""
        s0, s1 = self._pairs[k]
        cap = len(s0) + len(s1)
        n = int(self._alen[k])
        ops = self.moves(k, cap)
        A = np.full(n, ord("-"), dtype=np.uint8)
        B = np.full(n, ord("-"), dtype=np.uint8)
        A[ops != 2] = np.frombuffer(s0, dtype=np.uint8)
        B[ops != 0] = np.frombuffer(s1, dtype=np.uint8)
        return A.tobytes(), B.tobytes()

    def output_bytes(self, k: int) -> bytes:
        a, b = self.alignment(k)
        return b">1\n" + a + b"\n>2\n" + b


def psa_batch(pairs, M=2, X=-5, E=-2, O=-4, trace=False, packed=False, bits2=False) -> BatchResult:
    """Many independent pairs through tsta_psa_batch.  `pairs` is a list of (x, y) byte strings
    in any order; each pair is swapped like psa/psa.c:194-199.  bits2: ship the (ACGT) reads in the
    2-bit wire format (tsta_psa_batch_2bit)."""
    ordered = [order_pair(x, y) for x, y in pairs]
    swapped = np.array([len(y) > len(x) for x, y in pairs], dtype=bool)
    npairs = len(ordered)
    len0 = np.array([len(a) for a, _ in ordered], dtype=np.int32)
    len1 = np.array([len(b) for _, b in ordered], dtype=np.int32)
    off0 = np.zeros(npairs, dtype=np.int64)
    off1 = np.zeros(npairs, dtype=np.int64)
    if npairs:
        off0[1:] = np.cumsum(len0[:-1], dtype=np.int64)
        off1[1:] = np.cumsum(len1[:-1], dtype=np.int64)
    seqs0 = b"".join(a for a, _ in ordered) or b"\0"
    seqs1 = b"".join(b for _, b in ordered) or b"\0"
    if bits2:
        seqs0, seqs1 = pack_2bit(seqs0), pack_2bit(seqs1)
    return psa_batch_arrays(seqs0, off0, len0, seqs1, off1, len1, M, X, E, O, trace,
                            pairs=ordered, swapped=swapped, packed=packed, bits2=bits2)


def psa_batch_arrays(seqs0, off0, len0, seqs1, off1, len1, M=2, X=-5, E=-2, O=-4, trace=False,
                     pairs=None, swapped=None, out=None, packed=False, bits2=False) -> BatchResult:
    """Array form of tsta_psa_batch (what bench.py uses; no per-pair Python objects).
    `out` = (maxsorce, corner[, ops, ops_off, alen]) preallocated arrays (e.g. pinned) to write into.
    bits2: seqs0 / seqs1 are in the 2-bit wire format (pack_2bit), offsets and lengths in bases (tsta_psa_batch_2bit;
    traced results then come back packed)."""
    if bits2:
        packed = True
    npairs = len(len0)
    ms = out[0] if out else np.zeros(npairs, dtype=np.int32)
    corner = out[1] if out else np.zeros(npairs, dtype=np.int32)
    p = Params(M, X, E, O)
    kms = C.c_float()
    ops = ops_off = alen = None
    if trace and out and len(out) >= 5:
        ops, ops_off, alen = out[2], out[3], out[4]
    elif trace:
        cap = len0.astype(np.int64) + len1.astype(np.int64)
        if packed:
            cap = (cap + 3) & ~3                    # packed slots start at multiples of four moves
        ops_off = np.zeros(npairs, dtype=np.int64)
        if npairs:
            ops_off[1:] = np.cumsum(cap[:-1])
        ops = np.zeros(max(int(cap.sum()) // (4 if packed else 1), 1), dtype=np.uint8)
        alen = np.zeros(npairs, dtype=np.int32)
    b0 = seqs0 if isinstance(seqs0, bytes) else seqs0.ctypes.data
    b1 = seqs1 if isinstance(seqs1, bytes) else seqs1.ctypes.data
    if bits2:
        L = load()
        L.tsta_psa_batch_2bit.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _check(L.tsta_psa_batch_2bit(
            C.byref(p), npairs,
            C.cast(b0, C.c_void_p), off0.ctypes.data, len0.ctypes.data,
            C.cast(b1, C.c_void_p), off1.ctypes.data, len1.ctypes.data,
            int(trace), ms.ctypes.data, corner.ctypes.data,
            ops.ctypes.data if trace else None, ops_off.ctypes.data if trace else None,
            alen.ctypes.data if trace else None, C.byref(kms)))
    elif trace and packed:
        L = load()
        L.tsta_psa_batch_packed.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _check(L.tsta_psa_batch_packed(
            C.byref(p), npairs,
            C.cast(b0, C.c_void_p), off0.ctypes.data, len0.ctypes.data,
            C.cast(b1, C.c_void_p), off1.ctypes.data, len1.ctypes.data,
            ms.ctypes.data, corner.ctypes.data, ops.ctypes.data, ops_off.ctypes.data, alen.ctypes.data, C.byref(kms)))
    else:
        _check(load().tsta_psa_batch(
            C.byref(p), npairs,
            C.cast(b0, C.c_void_p), off0.ctypes.data, len0.ctypes.data,
            C.cast(b1, C.c_void_p), off1.ctypes.data, len1.ctypes.data,
            int(trace), ms.ctypes.data, corner.ctypes.data,
            ops.ctypes.data if trace else None, ops_off.ctypes.data if trace else None,
            alen.ctypes.data if trace else None, C.byref(kms)))
    return BatchResult(ms, corner, kms.value,
                       swapped if swapped is not None else np.zeros(npairs, dtype=bool),
                       ops, ops_off, alen, pairs, bool(trace and packed))


def pack_2bit(seq) -> np.ndarray:
    """ACGT bytes -> the 2-bit wire format of tsta_psa_batch_2bit (base i in bits 2(i & 3).. of byte i >> 2)."""
    a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.asarray(seq, dtype=np.uint8)
    code = np.zeros(256, dtype=np.uint8)
    code[ord("C")], code[ord("G")], code[ord("T")] = 1, 2, 3
    c = code[a]
    pad = (-len(c)) % 4
    if pad:
        c = np.concatenate([c, np.zeros(pad, dtype=np.uint8)])
    c = c.reshape(-1, 4)
    return (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)


# ------------------------------------------------------------------------------------------
# POA (TSTA_msa hot path)
# ------------------------------------------------------------------------------------------
class PoaStep(C.Structure):
    _fields_ = [("row", C.c_int32), ("col", C.c_int32), ("kind", C.c_uint8), ("pred", C.c_uint8),
                ("reserved", C.c_uint16)]


class Poa:
    """Device context of the sequence-to-graph DP (tsta_poa_*)."""

    def __init__(self):
        L = load()
        L.tsta_poa_create.argtypes = [C.POINTER(C.c_void_p)]
        L.tsta_poa_destroy.argtypes = [C.c_void_p]
        L.tsta_poa_fill.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.POINTER(C.c_float)]
        L.tsta_poa_walk.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        L.tsta_poa_debug_trace.argtypes = [C.c_void_p, C.c_void_p]
        self._h = C.c_void_p()
        _check(L.tsta_poa_create(C.byref(self._h)))
        self.nrows = self.len = 0

    def fill(self, row_ptr, pred_row, base, hneg1, read: bytes, M=2, X=-5, E=-2, O=-4, L=640):
        """Returns (lastsorce[nrows] int32, kernel_ms)."""
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
        pred_row = np.ascontiguousarray(pred_row, dtype=np.int32)
        base = np.ascontiguousarray(base, dtype=np.uint8)
        hneg1 = np.ascontiguousarray(hneg1, dtype=np.int32)
        self.nrows, self.len = len(base), len(read)
        last = np.zeros(self.nrows, dtype=np.int32)
        ms = C.c_float()
        p = Params(M, X, E, O)
        _check(load().tsta_poa_fill(self._h, C.byref(p), L, self.nrows, row_ptr.ctypes.data, pred_row.ctypes.data,
                                    base.ctypes.data, hneg1.ctypes.data, read, self.len, last.ctypes.data,
                                    C.byref(ms)))
        return last, ms.value

    NODE2 = np.dtype([("id", "<i4"), ("base_in", "<u4"), ("pre", "<i4", (2,))])
    NODE8 = np.dtype([("id", "<i4"), ("base_in", "<u4"), ("pre", "<i4", (8,))])

    def graph_update(self, nnodes: int, nodes) -> None:
        """tsta_poa_graph_update: `nodes` = iterable of (id, base, [predecessor node ids]) (a head: empty list)."""
        L = load()
        L.tsta_poa_graph_update.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        nodes = list(nodes)
        narrow = np.zeros(sum(1 for n in nodes if len(n[2]) <= 2), dtype=self.NODE2)
        wide = np.zeros(len(nodes) - len(narrow), dtype=self.NODE8)
        i2 = i8 = 0
        for nid, base, pre in nodes:
            rec = narrow if len(pre) <= 2 else wide
            k = i2 if len(pre) <= 2 else i8
            rec[k]["id"], rec[k]["base_in"] = nid, base | (len(pre) << 8)
            rec[k]["pre"][:] = -1
            rec[k]["pre"][:len(pre)] = pre
            if len(pre) <= 2:
                i2 += 1
            else:
                i8 += 1
        _check(L.tsta_poa_graph_update(self._h, nnodes, len(narrow), narrow.ctypes.data if len(narrow) else None,
                                       len(wide), wide.ctypes.data if len(wide) else None))

    def fill_resident(self, sort, read: bytes, M=2, X=-5, E=-2, O=-4, L=640, hneg1=None):
        """tsta_poa_fill_resident: sort[r] = node id of row r.  Returns (lastsorce[nrows] int32, kernel_ms)."""
        Lb = load()
        Lb.tsta_poa_fill_resident.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_char_p, C.c_int, C.c_void_p, C.POINTER(C.c_float)]
        sort = np.ascontiguousarray(sort, dtype=np.int32)
        h = None if hneg1 is None else np.ascontiguousarray(hneg1, dtype=np.int32)
        self.nrows, self.len = len(sort), len(read)
        last = np.zeros(self.nrows, dtype=np.int32)
        ms = C.c_float()
        p = Params(M, X, E, O)
        _check(Lb.tsta_poa_fill_resident(self._h, C.byref(p), L, self.nrows, sort.ctypes.data,
                                         None if h is None else h.ctypes.data, read, self.len, last.ctypes.data, C.byref(ms)))
        return last, ms.value

    def h2d_bytes(self) -> int:
        L = load()
        L.tsta_poa_h2d_bytes.argtypes = [C.c_void_p]
        L.tsta_poa_h2d_bytes.restype = C.c_longlong
        return int(L.tsta_poa_h2d_bytes(self._h))

    def walk(self, start_row: int):
        """Returns (steps as a structured numpy array, kernel_ms)."""
        path = np.zeros(self.nrows + self.len + 1, dtype=np.dtype(
            [("row", "<i4"), ("col", "<i4"), ("kind", "u1"), ("pred", "u1"), ("reserved", "<u2")]))
        n, ms = C.c_int(), C.c_float()
        _check(load().tsta_poa_walk(self._h, start_row, path.ctypes.data, C.byref(n), C.byref(ms)))
        return path[:n.value], ms.value

    def trace(self) -> np.ndarray:
        out = np.zeros((self.nrows, self.len), dtype=np.uint16)
        _check(load().tsta_poa_debug_trace(self._h, out.ctypes.data))
        return out

    def close(self):
        if self._h:
            load().tsta_poa_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
