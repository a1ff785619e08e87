"""Fused instance normalisation (+ReLU) for RAFT/GMA's feature encoder (row f-4 of SURVEY.md section 8: network glue
around the cuDNN convolutions).  Host-side mirror of `relu(nn.InstanceNorm2d(C)(x))` as models/raft/extractor.py:13-55
applies it; CUDA tensors only — the encoder falls back to the nn modules for CPU tensors (oracle / CPU tests).
Works on NCHW-contiguous and on channels_last tensors (the output keeps the input's memory format)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import _lib


def _is_cl(x: torch.Tensor) -> bool:
    return (x.shape[1] % 4 == 0 and x.shape[1] <= 1024 and not x.is_contiguous()
            and x.is_contiguous(memory_format=torch.channels_last))


class _InstNormHalfFn(Function):
    """channels-last fp16 input (a convolution output under autocast) -> fp32 output; the gradient goes back in fp16."""

    @staticmethod
    def forward(ctx, x, eps, relu):
        lib = _lib.load()
        B, C, H, W = x.shape
        y = torch.empty((B, C, H, W), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
        stats = torch.empty(B * C, 2, device=x.device, dtype=torch.float32)
        ws = torch.empty(lib.pcfa_instnorm_workspace_bytes(B, C, H, W), device=x.device, dtype=torch.uint8)
        _lib.check(lib.pcfa_instnorm_forward_h(_lib.ptr(x), _lib.ptr(y), _lib.ptr(stats), _lib.ptr(ws), B, C, H, W, float(eps), int(relu),
                                               _lib.stream()), "pcfa_instnorm_forward_h")
        ctx.save_for_backward(x, stats)
        ctx.relu = int(relu)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, stats = ctx.saved_tensors
        gy = gy.float().contiguous(memory_format=torch.channels_last)
        lib = _lib.load()
        B, C, H, W = x.shape
        gx = torch.empty_like(x)
        ws = torch.empty(lib.pcfa_instnorm_workspace_bytes(B, C, H, W), device=x.device, dtype=torch.uint8)
        _lib.check(lib.pcfa_instnorm_backward_h(_lib.ptr(x), _lib.ptr(gy), _lib.ptr(stats), _lib.ptr(gx), _lib.ptr(ws), B, C, H, W, ctx.relu,
                                                _lib.stream()), "pcfa_instnorm_backward_h")
        return gx, None, None


class _InstNormFn(Function):
    @staticmethod
    def forward(ctx, x, eps, relu):
        cl = _is_cl(x)
        if not cl:
            x = x.contiguous()
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("instance_norm: expected a CUDA float32 tensor (pcfa_b200 has no CPU path)")
        lib = _lib.load()
        B, C, H, W = x.shape
        y = torch.empty_like(x)                          # preserves the memory format
        stats = torch.empty(B * C, 2, device=x.device, dtype=torch.float32)
        ws = torch.empty(lib.pcfa_instnorm_workspace_bytes(B, C, H, W), device=x.device, dtype=torch.uint8)
        _lib.check(lib.pcfa_instnorm_forward(_lib.ptr(x), _lib.ptr(y), _lib.ptr(stats), _lib.ptr(ws), B, C, H, W,
                                             float(eps), int(relu), int(cl), _lib.stream()), "pcfa_instnorm_forward")
        ctx.save_for_backward(x, stats)
        ctx.relu, ctx.cl = int(relu), cl
        return y

    @staticmethod
    def backward(ctx, gy):
        x, stats = ctx.saved_tensors
        gy = gy.contiguous(memory_format=torch.channels_last) if ctx.cl else gy.contiguous()
        lib = _lib.load()
        B, C, H, W = x.shape
        gx = torch.empty_like(x)
        ws = torch.empty(lib.pcfa_instnorm_workspace_bytes(B, C, H, W), device=x.device, dtype=torch.uint8)
        _lib.check(lib.pcfa_instnorm_backward(_lib.ptr(x), _lib.ptr(gy), _lib.ptr(stats), _lib.ptr(gx), _lib.ptr(ws),
                                              B, C, H, W, ctx.relu, int(ctx.cl), _lib.stream()), "pcfa_instnorm_backward")
        return gx, None, None


def instance_norm(x: torch.Tensor, eps: float = 1e-5, relu: bool = False) -> torch.Tensor:
    """relu?(F.instance_norm(x, eps=eps)) for a CUDA 4-D tensor (no affine, no running statistics).  Half-precision
    inputs (convolution outputs under autocast, GMA's default) are normalised in fp32 and returned in fp32, which is
    what autocast makes of F.instance_norm as well."""
    if x.dtype == torch.float16 and x.is_cuda and x.dim() == 4 and _is_cl(x):
        return _InstNormHalfFn.apply(x, eps, relu)            # no conversion copies (csrc/instnorm.cu, *_h entry points)
    if x.dtype != torch.float32:
        x = x.float()
    return _InstNormFn.apply(x, eps, relu)


def fusable(norm: torch.nn.Module, x: torch.Tensor) -> bool:
    """True when `norm` is a plain nn.InstanceNorm2d that the fused kernel reproduces and x lives on a GPU."""
    return (isinstance(norm, torch.nn.InstanceNorm2d) and not norm.affine and not norm.track_running_stats
            and x.is_cuda and x.dtype in (torch.float32, torch.float16, torch.bfloat16) and x.dim() == 4)
