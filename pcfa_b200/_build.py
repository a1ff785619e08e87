"""Builds libpcfa_b200.so (the C-ABI library of include/pcfa_b200.h) in-tree with nvcc for sm_100a.

No torch headers are involved: the library is plain CUDA C++ with an extern "C" surface, loaded
through ctypes by pcfa_b200._lib.  `python -m pcfa_b200._build` rebuilds it.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "lib" / "obj"
LIB = LIBDIR / "libpcfa_b200.so"

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
              "-Xptxas", "-v"]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")
    return cand


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(ARCH_FLAGS + NVCC_FLAGS).encode())
    return h.hexdigest()


def sources():
    return sorted(CSRC.glob("*.cu"))


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every csrc/*.cu for sm_100a and link libpcfa_b200.so.  Returns the library path."""
    srcs = sources()
    deps = srcs + sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "pcfa_b200.h"]
    stamp = LIBDIR / "build.sha256"
    digest = _digest(deps)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text().strip() == digest:
        return LIB
    OBJDIR.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    env = dict(os.environ)
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper); give nvcc a plain host compiler
    host_cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else None

    def compile_one(src: Path):
        obj = OBJDIR / (src.stem + ".o")
        cmd = [nvcc, *ARCH_FLAGS, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if host_cxx:
            cmd[1:1] = ["-ccbin", host_cxx]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        (OBJDIR / (src.stem + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, *ARCH_FLAGS, "-shared", "-o", str(LIB), *map(str, objs)]
    if host_cxx:
        cmd[1:1] = ["-ccbin", host_cxx]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
