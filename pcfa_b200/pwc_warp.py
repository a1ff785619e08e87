"""PWCNet's backward warp (PWCDCNet.warp, models/PWCNet/PWCNet.py:166-206) as one fused op:
grid_sample(x) * (grid_sample(ones) >= 1e-4), gradients w.r.t. x and flow."""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import _lib


class PWCWarpFunction(Function):
    @staticmethod
    def forward(ctx, x, flo):
        lib = _lib.load()
        x, flo = x.contiguous(), flo.contiguous()
        _lib.require_cuda(x, flo, name="pwc_warp")
        B, Cc, H, W = x.shape
        out = torch.empty_like(x)
        _lib.check(lib.pcfa_pwc_warp_forward(_lib.ptr(x), _lib.ptr(flo), _lib.ptr(out), B, Cc, H, W,
                                             _lib.stream()), "pcfa_pwc_warp_forward")
        ctx.save_for_backward(x, flo)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        x, flo = ctx.saved_tensors
        gout = gout.contiguous()
        B, Cc, H, W = x.shape
        gx = torch.zeros_like(x)
        gflo = torch.empty_like(flo)
        _lib.check(lib.pcfa_pwc_warp_backward(_lib.ptr(x), _lib.ptr(flo), _lib.ptr(gout), _lib.ptr(gx),
                                              _lib.ptr(gflo), B, Cc, H, W, _lib.stream()),
                   "pcfa_pwc_warp_backward")
        return gx, gflo


def pwc_warp(x, flo):
    return PWCWarpFunction.apply(x, flo)
