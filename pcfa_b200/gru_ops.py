"""Fused element-wise halves of RAFT/GMA's convolutional GRU (models/raft/update.py:16-60) — row f-4 of SURVEY.md
section 8.  `gru_gates(zr, h) -> (z, rh)` and `gru_blend(z, q_pre, h) -> h_new` replace eight ATen launches per GRU step
forward and about ten backward with two each; CUDA fp32 NCHW-contiguous tensors only."""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import _lib


class _Gates(Function):
    @staticmethod
    def forward(ctx, zr, h):
        zr, h = zr.contiguous(), h.contiguous()
        _lib.require_cuda(zr, h, name="gru_gates")
        B, C = h.shape[0], h.shape[1]
        if zr.shape[1] != 2 * C or zr.shape[0] != B or zr.shape[2:] != h.shape[2:]:
            raise RuntimeError(f"gru_gates: zr {tuple(zr.shape)} does not match h {tuple(h.shape)}")
        n = h.numel() // B
        z, r, rh = torch.empty_like(h), torch.empty_like(h), torch.empty_like(h)
        lib = _lib.load()
        _lib.check(lib.pcfa_gru_gates_forward(_lib.ptr(zr), _lib.ptr(h), _lib.ptr(z), _lib.ptr(r), _lib.ptr(rh), B, n,
                                              _lib.stream()), "pcfa_gru_gates_forward")
        ctx.save_for_backward(z, r, h)
        return z, rh

    @staticmethod
    def backward(ctx, gz, grh):
        z, r, h = ctx.saved_tensors
        B = h.shape[0]
        n = h.numel() // B
        gz = None if gz is None else gz.contiguous()
        grh = None if grh is None else grh.contiguous()
        gzr = torch.empty(B, 2 * h.shape[1], *h.shape[2:], device=h.device, dtype=h.dtype)
        gh = torch.empty_like(h)
        lib = _lib.load()
        _lib.check(lib.pcfa_gru_gates_backward(_lib.ptr(z), _lib.ptr(r), _lib.ptr(h), _lib.ptr(gz), _lib.ptr(grh),
                                               _lib.ptr(gzr), _lib.ptr(gh), B, n, _lib.stream()), "pcfa_gru_gates_backward")
        return gzr, gh


class _Blend(Function):
    @staticmethod
    def forward(ctx, z, q_pre, h):
        z, q_pre, h = z.contiguous(), q_pre.contiguous(), h.contiguous()
        _lib.require_cuda(z, q_pre, h, name="gru_blend")
        if not (z.shape == q_pre.shape == h.shape):
            raise RuntimeError("gru_blend: shape mismatch")
        q, hn = torch.empty_like(h), torch.empty_like(h)
        lib = _lib.load()
        _lib.check(lib.pcfa_gru_blend_forward(_lib.ptr(z), _lib.ptr(q_pre), _lib.ptr(h), _lib.ptr(q), _lib.ptr(hn),
                                              h.numel(), _lib.stream()), "pcfa_gru_blend_forward")
        ctx.save_for_backward(z, q, h)
        return hn

    @staticmethod
    def backward(ctx, ghn):
        z, q, h = ctx.saved_tensors
        ghn = ghn.contiguous()
        gz, gq, gh = torch.empty_like(h), torch.empty_like(h), torch.empty_like(h)
        lib = _lib.load()
        _lib.check(lib.pcfa_gru_blend_backward(_lib.ptr(z), _lib.ptr(q), _lib.ptr(h), _lib.ptr(ghn), _lib.ptr(gz),
                                               _lib.ptr(gq), _lib.ptr(gh), h.numel(), _lib.stream()), "pcfa_gru_blend_backward")
        return gz, gq, gh


def gru_gates(zr: torch.Tensor, h: torch.Tensor):
    """(sigmoid(zr[:, :C]), sigmoid(zr[:, C:]) * h) for zr = [B, 2C, H, W] pre-activations."""
    return _Gates.apply(zr, h)


def gru_blend(z: torch.Tensor, q_pre: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    """(1 - z) * h + z * tanh(q_pre)."""
    return _Blend.apply(z, q_pre, h)


def usable(*ts: torch.Tensor) -> bool:
    return all(t.is_cuda and t.dtype == torch.float32 for t in ts) and not torch.is_autocast_enabled()
