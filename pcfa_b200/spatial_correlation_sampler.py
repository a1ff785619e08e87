"""Spatial correlation sampler on B200 — drop-in for the vendored package the reference's PWCNet
imports (models/PWCNet/cpu_spatial_correlation_sampler-0.3.0/Correlation_Module/
spatial_correlation_sampler/spatial_correlation_sampler.py:9-107):

    spatial_correlation_sample(input1, input2, kernel_size=1, patch_size=1, stride=1,
                               padding=0, dilation=1, dilation_patch=1)  -> [B, pH, pW, oH, oW]
    SpatialCorrelationSampler(...)(input1, input2)

Unlike the reference's default build (CPU only, setup.py:5) the inputs stay on the GPU and the
kernels run on the current stream (the reference's CUDA build uses the legacy default stream,
correlation_cuda_kernel.cu:266).  Semantics follow the CPU implementation (correlation.cpp:75-178).
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from . import _lib


def _params(kernel_size, patch_size, stride, padding, dilation, dilation_patch):
    kH, kW = _pair(kernel_size)
    pH, pW = _pair(patch_size)
    padH, padW = _pair(padding)
    dilH, dilW = _pair(dilation)
    dpH, dpW = _pair(dilation_patch)
    dH, dW = _pair(stride)
    return _lib.ScsParams(kH, kW, pH, pW, padH, padW, dilH, dilW, dpH, dpW, dH, dW)


class SpatialCorrelationSamplerFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1,
                dilation_patch=1, scale=1.0):
        lib = _lib.load()
        input1 = input1.contiguous()
        input2 = input2.contiguous()
        _lib.require_cuda(input1, input2, name="spatial_correlation_sample")
        if input1.shape != input2.shape or input1.dim() != 4:
            raise RuntimeError("spatial_correlation_sample: inputs must be 4-D and of equal shape")
        p = _params(kernel_size, patch_size, stride, padding, dilation, dilation_patch)
        B, Cc, iH, iW = input1.shape
        oH, oW = C.c_int(), C.c_int()
        _lib.check(lib.pcfa_scs_output_size(iH, iW, C.byref(p), C.byref(oH), C.byref(oW)),
                   "pcfa_scs_output_size")
        out = torch.empty((B, p.patchH, p.patchW, oH.value, oW.value), device=input1.device,
                          dtype=torch.float32)
        _lib.check(lib.pcfa_scs_forward(_lib.ptr(input1), _lib.ptr(input2), _lib.ptr(out), B, Cc, iH, iW,
                                        C.byref(p), float(scale), _lib.stream()), "pcfa_scs_forward")
        ctx.save_for_backward(input1, input2)
        ctx.p, ctx.scale = p, float(scale)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        lib = _lib.load()
        input1, input2 = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        B, Cc, iH, iW = input1.shape
        g1 = torch.empty_like(input1)
        g2 = torch.empty_like(input2)
        _lib.check(lib.pcfa_scs_backward(_lib.ptr(input1), _lib.ptr(input2), _lib.ptr(grad_output),
                                         _lib.ptr(g1), _lib.ptr(g2), B, Cc, iH, iW, C.byref(ctx.p),
                                         ctx.scale, _lib.stream()), "pcfa_scs_backward")
        return g1, g2, None, None, None, None, None, None, None


def spatial_correlation_sample(input1, input2, kernel_size=1, patch_size=1, stride=1, padding=0,
                               dilation=1, dilation_patch=1):
    return SpatialCorrelationSamplerFunction.apply(input1, input2, kernel_size, patch_size, stride,
                                                   padding, dilation, dilation_patch)


class SpatialCorrelationSampler(nn.Module):
    def __init__(self, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1, dilation_patch=1):
        super().__init__()
        self.kernel_size, self.patch_size, self.stride = kernel_size, patch_size, stride
        self.padding, self.dilation, self.dilation_patch = padding, dilation, dilation_patch

    def forward(self, input1, input2):
        return SpatialCorrelationSamplerFunction.apply(input1, input2, self.kernel_size, self.patch_size,
                                                       self.stride, self.padding, self.dilation,
                                                       self.dilation_patch)


def pwc_correlate(input1, input2):
    """PWCNet's `correlate` (models/PWCNet/PWCNet.py:45-58): 9x9 patch, /C folded into the kernel,
    output viewed as [B, 81, H, W]; no device hop."""
    out = SpatialCorrelationSamplerFunction.apply(input1, input2, 1, 9, 1, 0, 1, 1,
                                                  1.0 / input1.size(1))
    b, ph, pw, h, w = out.size()
    return out.view(b, ph * pw, h, w)
