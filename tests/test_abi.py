"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every
symbol include/tsta_cuda.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from tsta_b200 import build, api
    build.build_lib()
    return api.load()


def declared_symbols():
    txt = (ROOT / "include" / "tsta_cuda.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tsta_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in tsta_cuda.h but not exported"


def test_no_cpu_fallback_without_device(lib):
    from tsta_b200 import api
    try:
        n = api.device_count()
    except api.TstaError:
        n = 0
    if n > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.TstaError) as e:
        api.psa(b"ACGT", b"ACG")
    assert e.value.code == -2  # TSTA_ENODEV


def test_product_never_imports_oracle():
    for p in (ROOT / "tsta_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".c", ".h"}:
            txt = p.read_text()
            assert not re.search(r"(import\s+oracle|from\s+oracle|#include[^\n]*oracle|dlopen[^\n]*oracle|oracle[._]binding)", txt), p
