"""ctypes binding of the CPU oracle (oracle/psa_oracle.c, oracle/poa_oracle.c) and a runner
for the real reference binaries built by oracle/build_ref.sh.

TEST INFRASTRUCTURE ONLY.  Importers allowed by the project rules: tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing in
tsta_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from dataclasses import dataclass
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build() -> Path:
    """Compile the oracle shared library (gcc, a second or two)."""
    subprocess.run(["make", "-s", "-C", str(_HERE)], check=True)
    return _HERE / "_build" / "libtsta_oracle.so"


class _Result(C.Structure):
    _fields_ = [("maxsorce", C.c_int), ("corner", C.c_int),
                ("lastrow_fnv", C.c_uint64), ("alen", C.c_int64)]


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = _HERE / "_build" / "libtsta_oracle.so"
        srcs = [p for p in _HERE.glob("*.c")]
        if not so.exists() or any(p.stat().st_mtime > so.stat().st_mtime for p in srcs):
            build()
        _LIB = C.CDLL(str(so))
        _LIB.tsta_oracle_psa.restype = C.c_int
        _LIB.tsta_oracle_psa.argtypes = [
            C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
            C.POINTER(_Result), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]
        _LIB.tsta_oracle_psa_ref3.restype = C.c_int
        _LIB.tsta_oracle_psa_ref3.argtypes = [
            C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
            C.POINTER(_Result), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        _LIB.tsta_oracle_psa_scores.restype = C.c_int
        _LIB.tsta_oracle_psa_scores.argtypes = [
            C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
            C.POINTER(_Result), C.c_void_p]
        _LIB.tsta_oracle_free.argtypes = [C.c_void_p]
    return _LIB


@dataclass
class PsaResult:
    maxsorce: int
    corner: int
    lastrow_fnv: int
    a: bytes | None = None
    b: bytes | None = None
    lastrow: np.ndarray | None = None
    dirs: np.ndarray | None = None

    def output_bytes(self) -> bytes:
        """The bytes TSTA_psa writes to its -o file (reference psa/psa.c:512-515)."""
        return b">1\n" + self.a + b"\n>2\n" + self.b


def order_pair(x: bytes, y: bytes):
    """Reference psa/psa.c:194-199: the longer read becomes s0 (columns); ties keep order."""
    return (y, x) if len(y) > len(x) else (x, y)


def psa_scores(x: bytes, y: bytes, M=2, X=-5, E=-2, O=-4) -> PsaResult:
    """Score-only oracle for the full BASELINE sizes (no direction bookkeeping): maxsorce, corner,
    the whole last row."""
    s0, s1 = order_pair(x, y)
    n, m = len(s0), len(s1)
    r = _Result()
    lastrow = np.zeros(n, dtype=np.int32)
    lib().tsta_oracle_psa_scores(s0, n, s1, m, M, X, E, O, C.byref(r), lastrow.ctypes.data)
    return PsaResult(r.maxsorce, r.corner, r.lastrow_fnv, lastrow=lastrow)


def psa(x: bytes, y: bytes, M=2, X=-5, E=-2, O=-4, trace=True, want_lastrow=False,
        want_dirs=False, literal=False) -> PsaResult:
    """Run the oracle on reads x,y (any order; swapped like the reference does)."""
    s0, s1 = order_pair(x, y)
    n, m = len(s0), len(s1)
    r = _Result()
    a, b = C.c_void_p(), C.c_void_p()
    L = lib()
    if literal:
        rc = L.tsta_oracle_psa_ref3(s0, n, s1, m, M, X, E, O, C.byref(r), C.byref(a), C.byref(b))
        lastrow = dirs = None
    else:
        lastrow = np.zeros(n, dtype=np.int32) if want_lastrow else None
        dirs = np.zeros((m, n), dtype=np.uint8) if want_dirs else None
        rc = L.tsta_oracle_psa(
            s0, n, s1, m, M, X, E, O, int(trace), C.byref(r), C.byref(a), C.byref(b),
            lastrow.ctypes.data if lastrow is not None else None,
            dirs.ctypes.data if dirs is not None else None)
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    out = PsaResult(r.maxsorce, r.corner, r.lastrow_fnv, lastrow=lastrow, dirs=dirs)
    if a.value:
        out.a = C.string_at(a.value, r.alen)
        out.b = C.string_at(b.value, r.alen)
        L.tsta_oracle_free(a)
        L.tsta_oracle_free(b)
    return out


# ----------------------------------------------------------------------------------------
# the real reference (oracle/_ref/<isa>/TSTA_*), when it has been built
# ----------------------------------------------------------------------------------------
def _cpu_flags() -> set:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def ref_isa() -> str | None:
    """Best ISA build of the reference this host can execute, or None if not built."""
    flags = _cpu_flags()
    order = []
    if {"avx512f", "avx512bw", "avx512dq", "avx512vl"} <= flags:
        order.append("avx512")
    if "avx2" in flags:
        order.append("avx2")
    order.append("sse")
    for isa in order:
        if (_HERE / "_ref" / isa / "TSTA_psa").exists():
            return isa
    return None


def ref_bin(name: str, isa: str | None = None) -> Path | None:
    isa = isa or ref_isa()
    if isa is None:
        return None
    p = _HERE / "_ref" / isa / name
    return p if p.exists() else None


def write_fasta(path, reads, width: int = 0) -> None:
    with open(path, "wb") as f:
        for k, r in enumerate(reads):
            f.write(b">" + str(k + 1).encode() + b"\n")
            if width:
                for o in range(0, len(r), width):
                    f.write(r[o:o + width] + b"\n")
            else:
                f.write(r + b"\n")


def run_ref(name: str, args: list, timeout: float = 600.0, retries: int = 3, isa=None):
    """Run a reference binary under a timeout; retry on the known shutdown hang
    (thread-pool teardown race, reference pthreadpool/pthreadpool.c:89-120)."""
    exe = ref_bin(name, isa)
    if exe is None:
        raise FileNotFoundError("reference not built: run oracle/build_ref.sh")
    last = None
    for _ in range(retries):
        try:
            return subprocess.run([str(exe)] + [str(a) for a in args], capture_output=True,
                                  timeout=timeout, check=False)
        except subprocess.TimeoutExpired as e:  # pragma: no cover - rare
            last = e
    raise last


def ref_psa(x: bytes, y: bytes, M=2, X=-5, E=-2, O=-4, T=4, W=10, trace=True, isa=None,
            timeout: float = 600.0):
    """Run the real TSTA_psa / TSTA_psa_notrace; returns (maxsorce, output_file_bytes|None)."""
    with tempfile.TemporaryDirectory() as d:
        f1, f2, fo = os.path.join(d, "a.fa"), os.path.join(d, "b.fa"), os.path.join(d, "o.txt")
        write_fasta(f1, [x])
        write_fasta(f2, [y])
        args = ["-M", M, "-X", X, "-E", E, "-O", O, "-T", T, "-W", W, "-1", f1, "-2", f2]
        if trace:
            args += ["-o", fo]
        p = run_ref("TSTA_psa" if trace else "TSTA_psa_notrace", args, timeout=timeout, isa=isa)
        txt = p.stdout.decode()
        if "maxsorce=" not in txt:
            raise RuntimeError("reference failed: " + txt + p.stderr.decode())
        ms = int(txt.strip().split("maxsorce=")[1])
        out = open(fo, "rb").read() if trace else None
    return ms, out


# ----------------------------------------------------------------------------------------
# POA oracle (oracle/poa_oracle.c)
# ----------------------------------------------------------------------------------------
POA_STEP = np.dtype([("row", "<i4"), ("col", "<i4"), ("kind", "u1"), ("pred", "u1"), ("pad", "<u2")])


def poa_fill(row_ptr, pred_row, base, hneg1, read: bytes, M=2, X=-5, E=-2, O=-4, L=640):
    """Oracle DP of `read` against a CSR graph in row order.
    Returns (source, esource, fsource) int8 [nrows, len] and lastsorce int32 [nrows]."""
    Lb = lib()
    Lb.tsta_oracle_poa_fill.restype = C.c_int
    Lb.tsta_oracle_poa_fill.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
    pred_row = np.ascontiguousarray(pred_row, dtype=np.int32)
    base = np.ascontiguousarray(base, dtype=np.uint8)
    hneg1 = np.ascontiguousarray(hneg1, dtype=np.int32)
    nrows, ln = len(base), len(read)
    src = np.zeros((nrows, ln), dtype=np.int8)
    es = np.zeros((nrows, ln), dtype=np.int8)
    fs = np.zeros((nrows, ln), dtype=np.int8)
    last = np.zeros(nrows, dtype=np.int32)
    rc = Lb.tsta_oracle_poa_fill(nrows, row_ptr.ctypes.data, pred_row.ctypes.data, base.ctypes.data,
                                 hneg1.ctypes.data, read, ln, M, X, E, O, L,
                                 src.ctypes.data, es.ctypes.data, fs.ctypes.data, last.ctypes.data)
    if rc != 0:
        raise MemoryError("oracle POA fill failed")
    return src, es, fs, last


def poa_walk(row_ptr, pred_row, src, es, fs, start_row: int):
    """Oracle traceback walk (modifies a copy of `src` like the reference does)."""
    Lb = lib()
    Lb.tsta_oracle_poa_walk.restype = C.c_int
    Lb.tsta_oracle_poa_walk.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_void_p]
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
    pred_row = np.ascontiguousarray(pred_row, dtype=np.int32)
    nrows, ln = src.shape
    s2 = src.copy()
    path = np.zeros(nrows + ln + 1, dtype=POA_STEP)
    n = Lb.tsta_oracle_poa_walk(nrows, row_ptr.ctypes.data, pred_row.ctypes.data, ln, s2.ctypes.data,
                                np.ascontiguousarray(es).ctypes.data, np.ascontiguousarray(fs).ctypes.data,
                                start_row, path.ctypes.data)
    return path[:n]
