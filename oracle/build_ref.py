"""Builds the reference's OWN spatial-correlation-sampler CPU extension from the sources where they
lie under /root/reference (never copied into this repo) into oracle/_ref/.  TEST INFRASTRUCTURE ONLY.

Sources: models/PWCNet/cpu_spatial_correlation_sampler-0.3.0/Correlation_Module/{correlation.cpp,
correlation_sampler.cpp} — the build the reference ships by default (setup.py:5 CPU_ONLY = True).
The resulting torch extension `spatial_correlation_sampler_backend` validates the C restatement
(oracle.c) and is the "reference"-kind CPU baseline for the local-window correlation.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
SRC_DIR = Path("/root/reference/models/PWCNet/cpu_spatial_correlation_sampler-0.3.0/Correlation_Module")
NAME = "spatial_correlation_sampler_backend"


def build() -> Path | None:
    """Compile when the reference checkout is present; returns the .so path (or None)."""
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


def load():
    """Import the prebuilt extension from oracle/_ref (no compilation, no /root/reference access)."""
    so = REF_DIR / f"{NAME}.so"
    if not so.exists():
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules.setdefault(NAME, mod)
    return mod


if __name__ == "__main__":
    print(build())
