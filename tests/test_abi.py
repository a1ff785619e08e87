"""The C-ABI library loads on a CPU-only box and exports every symbol include/pcfa_b200.h declares;
argument validation (no compute) behaves as documented."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "pcfa_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcfa_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pcfa_b200 import _build, _lib
    _build.build()
    raw = C.CDLL(str(_lib.lib_path()))
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(raw, n), f"libpcfa_b200.so does not export {n}"
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"


def test_status_strings_and_argument_errors():
    from pcfa_b200 import _lib
    lib = _lib.load()
    assert lib.pcfa_abi_version() == 1
    assert lib.pcfa_status_string(0) == b"ok"
    assert b"workspace" in lib.pcfa_status_string(-4)
    # null pointers / bad sizes are rejected before any CUDA call
    assert lib.pcfa_corr_lookup_forward(None, None, None, 1, 8, 8, 4, 4, None) == -1
    assert lib.pcfa_corr_pyramid_forward(None, None, None, None, 0, 1, 8, 8, 8, 4, 0, None) == -1
    assert lib.pcfa_resample2d_forward(None, None, None, 1, 3, 8, 8, 8, 8, 1, 1, None) == -1
    assert lib.pcfa_instnorm_forward(None, None, None, None, 1, 4, 8, 8, 1e-5, 1, 0, None) == -1
    assert lib.pcfa_gru_gates_x_forward(None, None, None, None, None, None, None, 128, 128, 10, None) == -1
    assert lib.pcfa_cat_channels_last(None, None, 2, None, 10, None) == -1
    assert lib.pcfa_lbfgs_direction(None, None, None, None, None, None, None, None, 10, 100, 0, 0, None) == -1
    assert lib.pcfa_lbfgs_workspace_bytes() > 0 and lib.pcfa_instnorm_workspace_bytes(2, 64, 220, 512) > 0
    with pytest.raises(RuntimeError, match="status -1"):
        _lib.check(-1, "x")


def test_layout_and_size_helpers_match_reference_arithmetic():
    from pcfa_b200 import _lib
    from pcfa_b200.corr_block import pyramid_layout
    offs, hs, ws = pyramid_layout(1, 55, 128, 4)          # Sintel 440x1024 / 8  
    assert (hs, ws) == ([55, 27, 13, 6], [128, 64, 32, 16])
    n = 55 * 128
    assert offs == [0, n * 7040, n * (7040 + 27 * 64), n * (7040 + 27 * 64 + 13 * 32), n * (7040 + 27 * 64 + 13 * 32 + 96)]
    assert offs[-1] * 4 == 261_324_800                    # 261.3 MB per sample
    lib = _lib.load()
    p = _lib.ScsParams(1, 1, 9, 9, 0, 0, 1, 1, 1, 1, 1, 1)
    oh, ow = C.c_int(), C.c_int()
    assert lib.pcfa_scs_output_size(24, 80, C.byref(p), C.byref(oh), C.byref(ow)) == 0
    assert (oh.value, ow.value) == (24, 80)
    p = _lib.ScsParams(3, 3, 5, 5, 1, 1, 1, 1, 2, 2, 2, 2)
    lib.pcfa_scs_output_size(11, 12, C.byref(p), C.byref(oh), C.byref(ow))
    assert (oh.value, ow.value) == ((11 + 2 - 3) // 2 + 1, (12 + 2 - 3) // 2 + 1)
    oc = C.c_int()
    assert lib.pcfa_fn2corr_output_size(48, 160, 20, 1, 20, 1, 2, C.byref(oc), C.byref(oh), C.byref(ow)) == 0
    assert (oc.value, oh.value, ow.value) == (441, 48, 160)      # FlowNetC.py:26-31
    assert lib.pcfa_fn2corr_output_size(48, 160, 3, 1, 20, 1, 2, C.byref(oc), C.byref(oh), C.byref(ow)) == -1
    # occupancy bitmap of the sparse backward: one row per 32 queries, one bit per 32-cell chunk of every level
    chunks = sum(-(-(h * w) // 32) for h, w in zip(hs, ws))                 # 220 + 54 + 13 + 3
    assert chunks == 290
    assert lib.pcfa_corr_occupancy_bytes(1, 55, 128, 4) == 220 * (-(-chunks // 32)) * 4 == 8800
    assert lib.pcfa_corr_occupancy_bytes(8, 55, 128, 4) == 8 * 8800
    assert lib.pcfa_corr_occupancy_bytes(0, 55, 128, 4) == 0
    # argument errors of the occupancy / glue entry points (validated before any launch)
    assert lib.pcfa_corr_occupancy_mark(None, 1, None, 1, 55, 128, 4, 4, None) == -1
    assert lib.pcfa_add_rows_inplace(None, None, 1, 4, 4, None) == -1
    assert lib.pcfa_flow_step(None, None, None, 8, 0, None, None, 8, 1, 55, 128, None) == -1


def test_no_cpu_fallback_on_cpu_tensors():
    import torch
    from pcfa_b200.corr_block import CorrBlock
    from pcfa_b200.flownet2_ops import ChannelNorm
    x = torch.randn(1, 4, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        CorrBlock(x, x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ChannelNorm()(x)


def test_reference_binding_bodies_compile_against_the_header():
    """INTEGRATION.md section B is kept as compilable code (integration/reference_bindings.cpp): the bodies a maintainer
    would put into the reference's pybind11 modules.  Syntax-checked (g++ -fsyntax-only) against the torch headers and
    include/pcfa_b200.h, so the documented boundary cannot drift from the C ABI."""
    import shutil
    import subprocess
    import sysconfig
    from pathlib import Path
    from torch.utils import cpp_extension
    root = Path(__file__).resolve().parents[1]
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else shutil.which("g++")
    cuda_inc = Path("/usr/local/cuda/include")
    if cxx is None or not cuda_inc.exists():
        pytest.skip("no host compiler / CUDA headers")
    inc = cpp_extension.include_paths() + [str(cuda_inc), str(root / "include"), sysconfig.get_paths()["include"]]
    cmd = [cxx, "-std=c++17", "-fsyntax-only"] + [f"-I{p}" for p in inc] + [str(root / "integration" / "reference_bindings.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-4000:]
