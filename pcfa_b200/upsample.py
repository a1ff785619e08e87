"""Convex up-sampling of the flow (RAFT.upsample_flow, models/raft/raft.py:72-83; GMA models/gma/network.py:59-70) as one
fused forward kernel and two backward kernels (csrc/upsample.cu) instead of ~22 ATen launches and two layout conversions
of the 16 MB mask.  `convex_upsample(flow, mask, mask_scale)` takes the mask head's RAW output in either memory format
(channels-last is consumed in place; NCHW is converted once) and returns [N,2,8H,8W]."""
from __future__ import annotations

import torch

from . import _lib


class _ConvexUpsample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, mask, mask_scale):
        lib = _lib.load()
        if not (flow.is_cuda and mask.is_cuda):
            raise RuntimeError("convex_upsample: expected CUDA tensors (pcfa_b200 has no CPU path), got %s / %s" % (flow.device, mask.device))
        N, two, H, W = flow.shape
        if two != 2 or tuple(mask.shape) != (N, 576, H, W):
            raise ValueError("convex_upsample: flow [N,2,H,W] and mask [N,576,H,W] expected, got %s and %s"
                             % (tuple(flow.shape), tuple(mask.shape)))
        flow = flow.float().contiguous()
        mask = mask.float().contiguous(memory_format=torch.channels_last)
        up = torch.empty((N, 2, 8 * H, 8 * W), device=flow.device, dtype=torch.float32)
        _lib.check(lib.pcfa_convex_upsample_forward(_lib.ptr(flow), _lib.ptr(mask), _lib.ptr(up), N, H, W, float(mask_scale),
                                                    _lib.stream()), "pcfa_convex_upsample_forward")
        ctx.save_for_backward(flow, mask)
        ctx.mask_scale = float(mask_scale)
        return up

    @staticmethod
    def backward(ctx, gup):
        lib = _lib.load()
        flow, mask = ctx.saved_tensors
        N, _, H, W = flow.shape
        gup = gup.float().contiguous()
        gflow = torch.empty_like(flow)
        gmask = torch.empty_like(mask)                      # channels-last, what the mask head's backward consumes
        wsb = lib.pcfa_convex_upsample_workspace_bytes(N, H, W)
        ws = torch.empty(wsb, device=flow.device, dtype=torch.uint8)
        _lib.check(lib.pcfa_convex_upsample_backward(_lib.ptr(flow), _lib.ptr(mask), _lib.ptr(gup), _lib.ptr(gflow), _lib.ptr(gmask),
                                                     _lib.ptr(ws), wsb, N, H, W, ctx.mask_scale, _lib.stream()),
                   "pcfa_convex_upsample_backward")
        return gflow, gmask, None


def convex_upsample(flow: torch.Tensor, mask: torch.Tensor, mask_scale: float = 1.0) -> torch.Tensor:
    return _ConvexUpsample.apply(flow, mask, mask_scale)
