"""Builds the reference's OWN native extensions from the sources where they lie under /root/reference
(never copied into this repo) into oracle/_ref/.  TEST INFRASTRUCTURE ONLY.

1. CPU: models/PWCNet/cpu_spatial_correlation_sampler-0.3.0/Correlation_Module/{correlation.cpp,
   correlation_sampler.cpp} — the build the reference ships by default (setup.py:5 CPU_ONLY = True) →
   `spatial_correlation_sampler_backend` (validates oracle.c; "reference"-kind CPU baseline).
2. CUDA, compiled for sm_100a (nvcc cross-compiles without a GPU; they only RUN on the GPU box, as the
   GPU-side oracle of rows a5/a7/a8/a9 and as "the kernel to beat on the same box"):
   * `correlation_cuda`   models/FlowNet/correlation_package/{correlation_cuda.cc, correlation_cuda_kernel.cu}
   * `resample2d_cuda`    models/FlowNet/resample2d_package/{resample2d_cuda.cc, resample2d_kernel.cu}
   * `channelnorm_cuda`   models/FlowNet/channelnorm_package/{channelnorm_cuda.cc, channelnorm_kernel.cu}
   * `spatial_correlation_sampler_backend_cuda`   the sampler's CUDA build (correlation.cpp,
     correlation_sampler.cpp with -DUSE_CUDA, correlation_cuda_kernel.cu)
   torch 2.11 removed `Tensor::type()` as a dispatch key: the two FlowNet2 kernels that use it are
   patched ON A TEMPORARY COPY (`.type()` → `.scalar_type()` inside AT_DISPATCH, nothing else); the copy
   lives in a temp dir and is deleted after the build.  Only the .so files land in oracle/_ref/.
"""
from __future__ import annotations

import importlib.util
import os
import re
import shutil
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
REF = Path("/root/reference")
SRC_DIR = REF / "models/PWCNet/cpu_spatial_correlation_sampler-0.3.0/Correlation_Module"
NAME = "spatial_correlation_sampler_backend"

FN = REF / "models/FlowNet"
CUDA_EXTS = {
    "correlation_cuda": dict(dir=FN / "correlation_package", files=["correlation_cuda.cc", "correlation_cuda_kernel.cu"],
                             headers=["correlation_cuda_kernel.cuh"], defs=[]),
    "resample2d_cuda": dict(dir=FN / "resample2d_package", files=["resample2d_cuda.cc", "resample2d_kernel.cu"],
                            headers=["resample2d_kernel.cuh"], defs=[]),
    "channelnorm_cuda": dict(dir=FN / "channelnorm_package", files=["channelnorm_cuda.cc", "channelnorm_kernel.cu"],
                             headers=["channelnorm_kernel.cuh"], defs=[]),
    "spatial_correlation_sampler_backend_cuda": dict(dir=SRC_DIR, files=["correlation.cpp", "correlation_sampler.cpp",
                                                                           "correlation_cuda_kernel.cu"],
                                                     headers=[], defs=["-DUSE_CUDA"]),
}
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _patch(text: str) -> str:
    # AT_DISPATCH_*(x.type(), ...) → x.scalar_type(): the only edit, on the temporary copy
    return re.sub(r"(AT_DISPATCH_[A-Z_]+\(\s*\w+)\.type\(\)", r"\1.scalar_type()", text)


def build_cpu() -> Path | None:
    so = REF_DIR / f"{NAME}.so"
    if so.exists():
        return so
    if not SRC_DIR.exists():
        return None
    from torch.utils import cpp_extension
    REF_DIR.mkdir(exist_ok=True)
    os.environ.setdefault("CXX", "/usr/bin/g++")
    cpp_extension.load(name=NAME,
                       sources=[str(SRC_DIR / "correlation.cpp"), str(SRC_DIR / "correlation_sampler.cpp")],
                       extra_cflags=["-fopenmp", "-O3"],
                       extra_ldflags=["-L/usr/lib/gcc/x86_64-linux-gnu/13", "-lgomp"],
                       build_directory=str(REF_DIR), is_python_module=True, verbose=False)
    return so if so.exists() else None


def build_cuda(name: str, verbose: bool = False) -> Path | None:
    """One reference CUDA extension for sm_100a → oracle/_ref/<name>.so (None when the reference is absent)."""
    so = REF_DIR / f"{name}.so"
    if so.exists():
        return so
    spec = CUDA_EXTS[name]
    if not spec["dir"].exists():
        return None
    from torch.utils import cpp_extension
    REF_DIR.mkdir(exist_ok=True)
    os.environ.setdefault("CXX", "/usr/bin/g++")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    with tempfile.TemporaryDirectory(prefix="pcfa_ref_") as td:
        td = Path(td)
        src = td / "src"
        bld = td / "build"
        src.mkdir()
        bld.mkdir()
        for f in spec["files"] + spec["headers"]:
            (src / f).write_text(_patch((spec["dir"] / f).read_text()))
        cpp_extension.load(name=name, sources=[str(src / f) for f in spec["files"]],
                           extra_cflags=["-O3", *spec["defs"]],
                           extra_cuda_cflags=["-O3", *ARCH, "-lineinfo", *spec["defs"]],
                           build_directory=str(bld), is_python_module=True, verbose=verbose)
        shutil.copy(bld / f"{name}.so", so)
    return so


def build() -> Path | None:
    """Compile everything when the reference checkout is present; returns the CPU .so path (or None)."""
    so = build_cpu()
    for name in CUDA_EXTS:
        try:
            build_cuda(name)
        except Exception as e:                      # a failing GPU-oracle build must not hide the CPU one
            print(f"[oracle/build_ref] {name} not built: {str(e)[-400:]}", file=sys.stderr)
    return so


def _load(name: str):
    so = REF_DIR / f"{name}.so"
    if not so.exists():
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules.setdefault(name, mod)
    return mod


def load():
    """Import the prebuilt CPU sampler from oracle/_ref (no compilation, no /root/reference access)."""
    return _load(NAME)


def load_cuda(name: str):
    """Import a prebuilt reference CUDA extension from oracle/_ref; None when it was not built."""
    return _load(name)


if __name__ == "__main__":
    print(build())
    print(sorted(p.name for p in REF_DIR.glob("*.so")))
