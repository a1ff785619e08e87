"""FlowNet2's three custom operators on B200 — drop-ins for
  models/FlowNet/correlation_package/correlation.py:12-66      (CorrelationFunction, Correlation)
  models/FlowNet/resample2d_package/resample2d.py:12-56        (Resample2dFunction, Resample2d)
  models/FlowNet/channelnorm_package/channelnorm.py:11-45      (ChannelNormFunction, ChannelNorm)
and, below the Python layer, for the pybind modules `correlation_cuda`, `resample2d_cuda`,
`channelnorm_cuda` (same call signatures: caller-provided tensors are resized / filled).
"""
from __future__ import annotations

import ctypes as C

import torch
from torch.autograd import Function
from torch.nn.modules.module import Module

from . import _lib


# ------------------------------------------------------------------ backend-module equivalents
class correlation_cuda:
    """Namespace with the reference backend's forward/backward (correlation_cuda.cc:10-171)."""

    @staticmethod
    def output_size(H, W, pad_size, kernel_size, max_displacement, stride1, stride2):
        lib = _lib.load()
        oc, oh, ow = C.c_int(), C.c_int(), C.c_int()
        _lib.check(lib.pcfa_fn2corr_output_size(H, W, pad_size, kernel_size, max_displacement, stride1,
                                                stride2, C.byref(oc), C.byref(oh), C.byref(ow)),
                   "pcfa_fn2corr_output_size")
        return oc.value, oh.value, ow.value

    @staticmethod
    def forward(input1, input2, rbot1, rbot2, output, pad_size, kernel_size, max_displacement,
                stride1, stride2, corr_multiply):
        lib = _lib.load()
        _lib.require_cuda(input1, input2, name="correlation_cuda.forward")
        B, Cc, H, W = input1.shape
        oc, oh, ow = correlation_cuda.output_size(H, W, pad_size, kernel_size, max_displacement,
                                                  stride1, stride2)
        output.resize_(B, oc, oh, ow)      # rbot1 / rbot2 (padded NHWC copies) are never materialised
        _lib.check(lib.pcfa_fn2corr_forward(_lib.ptr(input1), _lib.ptr(input2), _lib.ptr(output), B, Cc,
                                            H, W, pad_size, kernel_size, max_displacement, stride1,
                                            stride2, _lib.stream()), "pcfa_fn2corr_forward")
        return 1

    @staticmethod
    def backward(input1, input2, rbot1, rbot2, grad_output, grad_input1, grad_input2, pad_size,
                 kernel_size, max_displacement, stride1, stride2, corr_multiply):
        lib = _lib.load()
        grad_output = grad_output.contiguous()
        _lib.require_cuda(input1, input2, grad_output, name="correlation_cuda.backward")
        B, Cc, H, W = input1.shape
        grad_input1.resize_(B, Cc, H, W)
        grad_input2.resize_(B, Cc, H, W)
        _lib.check(lib.pcfa_fn2corr_backward(_lib.ptr(input1), _lib.ptr(input2), _lib.ptr(grad_output),
                                             _lib.ptr(grad_input1), _lib.ptr(grad_input2), B, Cc, H, W,
                                             pad_size, kernel_size, max_displacement, stride1, stride2,
                                             _lib.stream()), "pcfa_fn2corr_backward")
        return 1


class resample2d_cuda:
    @staticmethod
    def forward(input1, input2, output, kernel_size, bilinear):
        lib = _lib.load()
        _lib.require_cuda(input1, input2, output, name="resample2d_cuda.forward")
        B, Cc, H, W = input1.shape
        _, _, oH, oW = input2.shape
        _lib.check(lib.pcfa_resample2d_forward(_lib.ptr(input1), _lib.ptr(input2), _lib.ptr(output), B,
                                               Cc, H, W, oH, oW, kernel_size, int(bool(bilinear)),
                                               _lib.stream()), "pcfa_resample2d_forward")
        return 1

    @staticmethod
    def backward(input1, input2, grad_output, grad_input1, grad_input2, kernel_size, bilinear):
        lib = _lib.load()
        _lib.require_cuda(input1, input2, grad_output, grad_input1, grad_input2,
                          name="resample2d_cuda.backward")
        B, Cc, H, W = input1.shape
        _, _, oH, oW = input2.shape
        _lib.check(lib.pcfa_resample2d_backward(_lib.ptr(input1), _lib.ptr(input2), _lib.ptr(grad_output),
                                                _lib.ptr(grad_input1), _lib.ptr(grad_input2), B, Cc, H,
                                                W, oH, oW, kernel_size, int(bool(bilinear)),
                                                _lib.stream()), "pcfa_resample2d_backward")
        return 1


class channelnorm_cuda:
    @staticmethod
    def forward(input1, output, norm_deg):
        lib = _lib.load()
        _lib.require_cuda(input1, output, name="channelnorm_cuda.forward")
        B, Cc, H, W = input1.shape
        _lib.check(lib.pcfa_channelnorm_forward(_lib.ptr(input1), _lib.ptr(output), B, Cc, H, W,
                                                int(norm_deg), _lib.stream()), "pcfa_channelnorm_forward")
        return 1

    @staticmethod
    def backward(input1, output, grad_output, grad_input1, norm_deg):
        lib = _lib.load()
        _lib.require_cuda(input1, output, grad_output, grad_input1, name="channelnorm_cuda.backward")
        B, Cc, H, W = input1.shape
        _lib.check(lib.pcfa_channelnorm_backward(_lib.ptr(input1), _lib.ptr(output), _lib.ptr(grad_output),
                                                 _lib.ptr(grad_input1), B, Cc, H, W, int(norm_deg),
                                                 _lib.stream()), "pcfa_channelnorm_backward")
        return 1


# ------------------------------------------------------------------ autograd Functions + Modules
class CorrelationFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, pad_size=3, kernel_size=3, max_displacement=20, stride1=1,
                stride2=2, corr_multiply=1):
        input1, input2 = input1.contiguous(), input2.contiguous()
        ctx.save_for_backward(input1, input2)
        ctx.cfg = (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)
        output = input1.new_empty(0)
        correlation_cuda.forward(input1, input2, None, None, output, *ctx.cfg)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        g1, g2 = input1.new_empty(0), input1.new_empty(0)
        correlation_cuda.backward(input1, input2, None, None, grad_output, g1, g2, *ctx.cfg)
        return g1, g2, None, None, None, None, None, None


class Correlation(Module):
    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1):
        super().__init__()
        self.pad_size, self.kernel_size, self.max_displacement = pad_size, kernel_size, max_displacement
        self.stride1, self.stride2, self.corr_multiply = stride1, stride2, corr_multiply

    def forward(self, input1, input2):
        return CorrelationFunction.apply(input1, input2, self.pad_size, self.kernel_size,
                                         self.max_displacement, self.stride1, self.stride2,
                                         self.corr_multiply)


class Resample2dFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, kernel_size=1, bilinear=True):
        assert input1.is_contiguous()
        assert input2.is_contiguous()
        ctx.save_for_backward(input1, input2)
        ctx.kernel_size, ctx.bilinear = kernel_size, bilinear
        _, d, _, _ = input1.size()
        b, _, h, w = input2.size()
        output = input1.new_empty((b, d, h, w))
        resample2d_cuda.forward(input1, input2, output, kernel_size, bilinear)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        grad_output = grad_output.contiguous()
        input1, input2 = ctx.saved_tensors
        grad_input1 = torch.zeros_like(input1)
        grad_input2 = torch.empty_like(input2)
        resample2d_cuda.backward(input1, input2, grad_output, grad_input1, grad_input2,
                                 ctx.kernel_size, ctx.bilinear)
        return grad_input1, grad_input2, None, None


class Resample2d(Module):
    def __init__(self, kernel_size=1, bilinear=True):
        super().__init__()
        self.kernel_size, self.bilinear = kernel_size, bilinear

    def forward(self, input1, input2):
        return Resample2dFunction.apply(input1.contiguous(), input2.contiguous(), self.kernel_size,
                                        self.bilinear)


class ChannelNormFunction(Function):
    @staticmethod
    def forward(ctx, input1, norm_deg=2):
        assert input1.is_contiguous()
        b, _, h, w = input1.size()
        output = input1.new_empty((b, 1, h, w))
        channelnorm_cuda.forward(input1, output, norm_deg)
        ctx.save_for_backward(input1, output)
        ctx.norm_deg = norm_deg
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input1, output = ctx.saved_tensors
        grad_input1 = torch.empty_like(input1)
        channelnorm_cuda.backward(input1, output, grad_output.contiguous(), grad_input1, ctx.norm_deg)
        return grad_input1, None


class ChannelNorm(Module):
    def __init__(self, norm_deg=2):
        super().__init__()
        self.norm_deg = norm_deg

    def forward(self, input1):
        return ChannelNormFunction.apply(input1.contiguous(), self.norm_deg)
