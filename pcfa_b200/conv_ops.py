"""Frozen convolution + bias + ReLU as ONE autograd node (row f-4 glue): cuDNN convolution without bias, the fused in-place
epilogue y = relu?(x + b) (csrc/bias_act.cu), and a backward that is the ReLU mask plus cuDNN's data gradient.  torch's
own path is three forward launches (convolution, broadcasting bias add, clamp) and three autograd nodes.  Only used for
frozen weights (the attack never trains the network, attack_PCFA.py:45-46); anything else takes the stock modules."""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F
from torch.autograd import Function

from . import _lib

_CL = torch.channels_last
_DUMMY = {}
_ENABLED = os.environ.get("PCFA_CONV_ACT", "1") != "0"       # 0: stock convolution + bias + ReLU modules (A/B measurements)


def _is_cl(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.is_contiguous(memory_format=_CL) and not t.is_contiguous()


def _dummy_like(shape, dtype, device, cl):
    key = (tuple(shape), dtype, device, cl)
    d = _DUMMY.get(key)
    if d is None:                                   # convolution_backward reads only its sizes / strides for the data gradient
        d = torch.empty(tuple(shape), dtype=dtype, device=device, memory_format=_CL if cl else torch.contiguous_format)
        _DUMMY[key] = d
    return d


class _ConvBiasAct(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, dilation, groups, relu, slope=0.0):
        lib = _lib.load()
        y = F.conv2d(x, weight, None, stride, padding, dilation, groups)
        cl = _is_cl(y)
        if not (cl or y.is_contiguous()):
            y = y.contiguous()
        C = y.shape[1]
        inner = 1 if cl else y.shape[2] * y.shape[3]
        st = lib.pcfa_bias_act_forward(_lib.ptr(y), _lib.ptr(bias), y.numel(), C, inner, int(relu), float(slope), 0 if y.dtype == torch.float32 else 1,
                                       _lib.stream())
        if st == -1:                                 # PCFA_E_BADARG: shape the vector kernels do not take (e.g. 2 output channels)
            y.add_(bias.view(1, -1, 1, 1))
            if relu:
                y = F.leaky_relu_(y, slope) if slope else y.relu_()
        else:
            _lib.check(st, "pcfa_bias_act_forward")
        ctx.save_for_backward(weight, y if relu else None)
        ctx.meta = (tuple(x.shape), x.dtype, x.device, _is_cl(x) or cl, stride, padding, dilation, groups, bool(relu), float(slope))
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        weight, y = ctx.saved_tensors
        shape, dtype, device, cl, stride, padding, dilation, groups, relu, slope = ctx.meta
        if relu:
            g = g.contiguous(memory_format=_CL) if _is_cl(y) else g.contiguous()
            if g.dtype != y.dtype:
                g = g.to(y.dtype)
            gx = torch.empty_like(y)
            st = lib.pcfa_relu_mask_backward(_lib.ptr(y), _lib.ptr(g), _lib.ptr(gx), y.numel(), slope, 0 if y.dtype == torch.float32 else 1,
                                             _lib.stream())
            if st == -1:
                gx = torch.where(y > 0, g, g * slope)
            else:
                _lib.check(st, "pcfa_relu_mask_backward")
            g = gx
        elif g.dtype != weight.dtype:
            g = g.to(weight.dtype)
        gin = torch.ops.aten.convolution_backward(g, _dummy_like(shape, dtype, device, cl), weight, None, stride, padding, dilation,
                                                  False, (0, 0), groups, (True, False, False))[0]
        return gin, None, None, None, None, None, None, None, None


def conv_act(conv: torch.nn.Conv2d, x: torch.Tensor, relu: bool, weight=None, bias=None, tag: str = "_pcfa_w16", slope: float = 0.0):
    """act?(conv(x)) with `conv`'s geometry (act = ReLU, or LeakyReLU(slope) for slope > 0); `weight` / `bias` override the
    module's (e.g. batch-norm-folded copies)."""
    w = conv.weight if weight is None else weight
    b = conv.bias if bias is None else bias
    frozen = not (w.requires_grad or (b is not None and b.requires_grad))
    if x.is_cuda and frozen and b is not None and conv.padding_mode == "zeros" and x.dim() == 4 and _ENABLED:
        from .networks.amp import amp_half_active, half_params
        if amp_half_active(x):
            w, b = half_params(conv, w, b, tag)
            if x.dtype != torch.float16:
                x = x.to(torch.float16)
        if x.dtype == w.dtype and x.dtype in (torch.float32, torch.float16):
            return _ConvBiasAct.apply(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups, relu, slope)
    y = F.conv2d(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
    return (F.leaky_relu(y, slope) if slope else F.relu(y)) if relu else y


class ConvLeakyReLU(torch.nn.Sequential):
    """nn.Sequential(Conv2d, LeakyReLU(slope)) with the same children and state-dict keys (PWCNet.py:23-28, FlowNet's
    submodules.py conv()), evaluated through conv_act when the weights are frozen and the input is a CUDA tensor."""

    def forward(self, x):
        return conv_act(self[0], x, True, slope=float(self[1].negative_slope))
