"""In-tree build of the CUDA C-ABI library and the C host drivers (nvcc + gcc, no GPU needed).

    python -m tsta_b200.build            # build everything that is out of date
    python -m tsta_b200.build --force

Outputs (git-ignored, but they travel to the GPU box with the snapshot):
    tsta_b200/_build/libtsta_cuda.so     CUDA kernels + C ABI (include/tsta_cuda.h), sm_100a only
    tsta_b200/_build/TSTA_psa, TSTA_psa_notrace, TSTA_msa    C command-line drivers
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
HOST = PKG / "host"
OUT = PKG / "_build"
LIB = OUT / "libtsta_cuda.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; tsta_b200 needs the CUDA toolkit to build (no CPU fallback)")


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    OUT.mkdir(exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    deps = srcs + sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "tsta_cuda.h"]
    if force or _newer(LIB, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-I", str(ROOT / "include"), "-o", str(LIB)] + [str(s) for s in srcs]
        subprocess.run(cmd, check=True)
    return LIB


def build_variant(name: str, defs) -> Path:
    """Experiment builds (tools/): the same library with -D overrides, as _build/libtsta_cuda_<name>.so;
    select it at run time with TSTA_LIB=<path>."""
    OUT.mkdir(exist_ok=True)
    out = OUT / f"libtsta_cuda_{name}.so"
    srcs = sorted(CSRC.glob("*.cu"))
    cmd = [_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defs] + ["-I", str(ROOT / "include"), "-o", str(out)] + [str(s) for s in srcs]
    subprocess.run(cmd, check=True)
    return out


def build_host(force: bool = False) -> list:
    """C drivers (the reference's three executables) linked against the C ABI."""
    OUT.mkdir(exist_ok=True)
    built = []
    if not HOST.exists():
        return built
    common = sorted(p for p in HOST.glob("*.c") if not p.name.startswith("tsta_"))
    hdrs = sorted(HOST.glob("*.h")) + [ROOT / "include" / "tsta_cuda.h"]
    targets = {
        "TSTA_psa": (["tsta_psa_main.c"], ["-DTRACE"]),
        "TSTA_psa_notrace": (["tsta_psa_main.c"], []),
        "TSTA_msa": (["tsta_msa_main.c"], []),
        "TSTA_psa_batch": (["tsta_psa_batch_main.c"], []),
        "tsta_fasta_dump": (["tsta_fasta_dump_main.c"], []),
    }
    for name, (mains, defs) in targets.items():
        srcs = [HOST / m for m in mains]
        if not all(s.exists() for s in srcs):
            continue
        exe = OUT / name
        if force or _newer(exe, srcs + common + hdrs + [LIB]):
            cmd = ["gcc", "-O2", "-g", "-std=gnu11", "-Wall"] + defs + \
                  ["-I", str(ROOT / "include"), "-I", str(HOST), "-o", str(exe)] + \
                  [str(s) for s in srcs + common] + \
                  ["-L", str(OUT), "-ltsta_cuda", "-Wl,-rpath,$ORIGIN", "-lz", "-lm", "-lstdc++", "-ldl", "-lpthread", "-lrt"]
            subprocess.run(cmd, check=True)
        built.append(exe)
    return built


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_lib(force, verbose)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", *(p.name for p in sorted(OUT.iterdir())))
