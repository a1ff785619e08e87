"""ctypes binding of libpcfa_b200.so — the only bridge between the PyTorch host code and the CUDA
kernels.  Signatures mirror include/pcfa_b200.h one to one.  There is no CPU fallback: if the
library is missing or a tensor is not a CUDA tensor the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_LIB_PATH = Path(os.environ.get("PCFA_LIB") or (Path(__file__).resolve().parent / "lib" / "libpcfa_b200.so"))   # PCFA_LIB: experiments
_lib = None
_raw = None
_profile = None        # None, or dict: entry point name -> list of (start_event, end_event)

c_fp = C.c_void_p      # device pointers travel as void*
c_i = C.c_int
c_i64 = C.c_int64
c_f = C.c_float
c_d = C.c_double


class ScsParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("kH", "kW", "patchH", "patchW", "padH", "padW", "dilH", "dilW",
                                       "dilPatchH", "dilPatchW", "dH", "dW")]


# name -> (restype, argtypes); every symbol declared in include/pcfa_b200.h
SIGNATURES = {
    "pcfa_abi_version": (c_i, []),
    "pcfa_status_string": (C.c_char_p, [c_i]),
    "pcfa_launch_count": (c_i64, []),
    "pcfa_corr_pyramid_layout": (c_i, [c_i, c_i, c_i, c_i, C.POINTER(c_i64), C.POINTER(c_i), C.POINTER(c_i)]),
    "pcfa_corr_pyramid_workspace_bytes": (c_i64, [c_i, c_i, c_i, c_i, c_i]),
    "pcfa_corr_pyramid_forward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_i64, c_i, c_i, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_corr_pyramid_backward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_i64, c_i, c_i, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_corr_lookup_forward": (c_i, [c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_corr_lookup_backward": (c_i, [c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_corr_lookup_forward_cl": (c_i, [c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_corr_lookup_backward_cl": (c_i, [c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_corr_occupancy_bytes": (c_i64, [c_i, c_i, c_i, c_i]),
    "pcfa_corr_occupancy_mark": (c_i, [c_fp, c_i, c_fp, c_i, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_corr_pyramid_backward_occ": (c_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_i64, c_i, c_i, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_scs_output_size": (c_i, [c_i, c_i, C.POINTER(ScsParams), C.POINTER(c_i), C.POINTER(c_i)]),
    "pcfa_scs_forward": (c_i, [c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, C.POINTER(ScsParams), c_f, c_fp]),
    "pcfa_scs_backward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, C.POINTER(ScsParams), c_f, c_fp]),
    "pcfa_fn2corr_output_size": (c_i, [c_i, c_i, c_i, c_i, c_i, c_i, c_i, C.POINTER(c_i), C.POINTER(c_i), C.POINTER(c_i)]),
    "pcfa_fn2corr_forward": (c_i, [c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_fn2corr_backward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_resample2d_forward": (c_i, [c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_resample2d_backward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_channelnorm_forward": (c_i, [c_fp, c_fp, c_i, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_channelnorm_backward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_pwc_warp_forward": (c_i, [c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_pwc_warp_backward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_box_forward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_i, c_i, c_i64, c_f, c_f, c_fp]),
    "pcfa_box_backward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i64, c_f, c_f, c_fp]),
    "pcfa_box_delta_backward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_i, c_i, c_i64, c_f, c_fp]),
    "pcfa_sumsq_partials": (c_i, [c_fp, c_i64, c_fp, c_fp]),
    "pcfa_objective_loss": (c_i, [c_fp, c_fp, c_fp, c_fp, c_f, c_f, c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i,
                                  c_i, c_i, c_i, c_i, c_d, c_f, c_f, c_fp]),
    "pcfa_objective_workspace_bytes": (c_i64, []),
    "pcfa_instnorm_workspace_bytes": (c_i64, [c_i, c_i, c_i, c_i]),
    "pcfa_instnorm_forward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, c_f, c_i, c_i, c_fp]),
    "pcfa_instnorm_forward_h": (c_i, [c_fp, c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, c_f, c_i, c_fp]),
    "pcfa_instnorm_backward_h": (c_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_gru_gates_forward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_i, c_i64, c_i, c_fp]),
    "pcfa_gru_gates_backward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_i, c_i64, c_i, c_fp]),
    "pcfa_gru_blend_forward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_i64, c_fp]),
    "pcfa_gru_blend_backward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_i64, c_fp]),
    "pcfa_gru_gates_x_forward": (c_i, [c_fp] * 7 + [c_i, c_i, c_i64, c_fp]),
    "pcfa_gru_gates_x_backward": (c_i, [c_fp] * 7 + [c_i, c_i, c_i64, c_fp]),
    "pcfa_gru_blend_x_forward": (c_i, [c_fp] * 8 + [c_i, c_i, c_i64, c_fp]),
    "pcfa_gru_blend_x_backward": (c_i, [c_fp] * 8 + [c_i, c_i, c_i64, c_fp]),
    "pcfa_gru_gates_x_backward_acc": (c_i, [c_fp] * 8 + [c_i, c_i, c_i, c_i64, c_fp]),
    "pcfa_gru_blend_x_backward_acc": (c_i, [c_fp] * 10 + [c_i, c_i, c_i, c_i64, c_fp]),
    "pcfa_gru_step_combine": (c_i, [c_fp] * 8 + [c_i, c_i, c_i64, c_fp]),
    "pcfa_bias_act_forward": (c_i, [c_fp, c_fp, c_i64, c_i, c_i64, c_i, c_f, c_i, c_fp]),
    "pcfa_relu_mask_backward": (c_i, [c_fp, c_fp, c_fp, c_i64, c_f, c_i, c_fp]),
    "pcfa_relu_mask_backward_rows": (c_i, [c_fp, c_fp, c_fp, c_i64, c_i, c_i64, c_f, c_i, c_fp]),
    "pcfa_add_rows_inplace": (c_i, [c_fp, c_fp, c_i64, c_i, c_i64, c_fp]),
    "pcfa_add_rows": (c_i, [c_fp, c_fp, c_fp, c_i64, c_i, c_i64, c_fp]),
    "pcfa_relu_mask2_backward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_i64, c_fp]),
    "pcfa_add_relu_forward": (c_i, [c_fp, c_fp, c_fp, c_i64, c_i, c_fp]),
    "pcfa_flow_step": (c_i, [c_fp, c_fp, c_fp, c_i, c_i, c_fp, c_fp, c_i, c_i, c_i, c_i, c_fp]),
    "pcfa_cat2_channels_last_h": (c_i, [c_fp, c_fp, c_fp, c_i, c_i, c_i64, c_fp]),
    "pcfa_gru_gates_x_forward_h": (c_i, [c_fp] * 7 + [c_i, c_i, c_i64, c_fp]),
    "pcfa_gru_blend_x_forward_h": (c_i, [c_fp] * 8 + [c_i, c_i, c_i64, c_fp]),
    "pcfa_gru_gates_x_backward_acc_h": (c_i, [c_fp] * 8 + [c_i, c_i, c_i, c_i64, c_fp]),
    "pcfa_gru_blend_x_backward_acc_h": (c_i, [c_fp] * 10 + [c_i, c_i, c_i, c_i64, c_fp]),
    "pcfa_gru_step_combine_h": (c_i, [c_fp] * 8 + [c_i, c_i, c_i64, c_fp]),
    "pcfa_lbfgs_workspace_bytes": (c_i64, []),
    "pcfa_lbfgs_update_history": (c_i, [c_fp, c_fp, c_fp, c_f, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_i64, c_i, c_fp]),
    "pcfa_lbfgs_compact_workspace_bytes": (c_i64, [c_i]),
    "pcfa_lbfgs_direction_compact": (c_i, [c_fp] * 5 + [c_fp, c_fp, c_fp, c_f, c_f, c_fp, c_fp, c_i64, c_i, c_fp]),
    "pcfa_lbfgs_direction_step": (c_i, [c_fp] * 6 + [c_fp, c_fp, c_f, c_f, c_fp, c_fp, c_i64, c_i, c_fp]),
    "pcfa_lbfgs_store_pair": (c_i, [c_fp, c_fp, c_fp, c_f, c_fp, c_fp, c_fp, c_fp, c_i64, c_fp]),
    "pcfa_lbfgs_direction": (c_i, [c_fp] * 8 + [c_i64, c_i, c_i, c_i, c_fp]),
    "pcfa_cat_channels_last": (c_i, [c_fp, c_fp, c_i, c_fp, c_i64, c_fp]),
    "pcfa_cat_channels_last_pad": (c_i, [c_fp, c_fp, c_i, c_fp, c_i64, c_i, c_fp]),
    "pcfa_softmax_rows_f16_forward": (c_i, [c_fp, c_fp, c_i64, c_i, c_fp]),
    "pcfa_softmax_rows_f16_backward": (c_i, [c_fp, c_fp, c_fp, c_i64, c_i, c_fp]),
    "pcfa_convex_upsample_workspace_bytes": (c_i64, [c_i, c_i, c_i]),
    "pcfa_convex_upsample_forward": (c_i, [c_fp, c_fp, c_fp, c_i, c_i, c_i, c_f, c_fp]),
    "pcfa_convex_upsample_backward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_i64, c_i, c_i, c_i, c_f, c_fp]),
    "pcfa_instnorm_backward": (c_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, c_i, c_i, c_fp]),
}


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load the shared library (once) and attach argtypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing — build it with `python -m pcfa_b200._build` "
            "(or __graft_entry__.build()).  pcfa_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.pcfa_abi_version() != 1:
        raise RuntimeError("libpcfa_b200.so ABI version mismatch — rebuild")
    global _raw
    _raw = lib
    _lib = _LibProxy()
    apply_determinism_from_env()
    return _lib


def deterministic() -> bool:
    return os.environ.get("PCFA_DETERMINISTIC", "0") not in ("", "0")


def apply_determinism_from_env():
    """PCFA_DETERMINISTIC=1 (read once, when the library is loaded): bit-reproducible closures.  The library side cuts the
    cost-volume backward's work shares at unit boundaries (one reduce-add per output element, csrc/corr_allpairs_bwd_tc.cu);
    this side pins cuDNN to deterministic algorithms without autotuning.  The lookup scatters are already deterministic
    (unique addresses per launch, launches stream-ordered); the warp / resample2d image gradients (PWCNet, FlowNet2) still
    use floating-point RED.ADD and are NOT covered."""
    if deterministic():
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False


class _LibProxy:
    """Attribute access returns the raw ctypes function, or — while profiling is enabled
    (pcfa_b200.profiling) — a wrapper that brackets the call with CUDA events on the current stream."""

    def __getattr__(self, name):
        fn = getattr(_raw, name)
        if _profile is None or not name.startswith("pcfa_") or fn.restype is not c_i:
            return fn

        def timed(*args):
            global _bytes_hint
            hint, _bytes_hint = _bytes_hint, None
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            # queue filler: a ~100 us device spin keeps the GPU busy while the host marshals the ctypes call and
            # enqueues event, launches and event — otherwise the bracket of a short entry point (two 5 us launches)
            # times the host's enqueue latency, not the kernels
            torch.cuda._sleep(200000)
            e0.record()
            r = fn(*args)
            e1.record()
            _profile.setdefault(name, []).append((e0, e1, args if hint is None else ("bytes", hint)))
            return r
        return timed


_bytes_hint = None


def hint_bytes(n: int) -> None:
    """Algorithmic bytes of the NEXT profiled entry-point call, for callers whose arguments do not carry the shape
    (pointer arrays).  Ignored when profiling is off."""
    global _bytes_hint
    if _profile is not None:
        _bytes_hint = int(n)


def event_overhead_us(n: int = 30) -> float:
    """Median duration of an EMPTY event bracket behind the same device-side spin the profiled calls use: the part of
    every bracketed measurement that is event bookkeeping, not kernel time (subtracted in profiling.kernel_table)."""
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200000)
        e0.record()
        e1.record()
        ts.append((e0, e1))
    torch.cuda.synchronize()
    v = sorted(a.elapsed_time(b) * 1e3 for a, b in ts)
    return v[len(v) // 2]


def set_profile(store):
    """Enable (dict) or disable (None) per-call CUDA-event timing of the C-ABI entry points."""
    global _profile
    _profile = store


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().pcfa_status_string(status).decode()
        raise RuntimeError(f"{what or 'pcfa_b200'} failed with status {status}: {msg}")


def ptr(t):
    """Device pointer of a tensor (None → NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors, name="pcfa_b200 operator"):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(f"{name}: expected CUDA tensors (pcfa_b200 has no CPU path), got {t.device}")
        if t.dtype != torch.float32:
            raise RuntimeError(f"{name}: expected float32 tensors, got {t.dtype}")
        if not t.is_contiguous():
            raise RuntimeError(f"{name}: expected contiguous tensors")


def launch_count() -> int:
    return int(load().pcfa_launch_count())
