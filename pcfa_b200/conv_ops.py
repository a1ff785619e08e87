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
_CUDNN_FUSED = os.environ.get("PCFA_CONV_FUSED", "1") != "0"   # 0: convolution + pcfa_bias_act_forward instead of cuDNN's fused epilogue
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


def _channel_slice_ld(g: torch.Tensor, y: torch.Tensor) -> int:
    """Pixel stride of g if g is a channel slice [:, a:b] of a wider channels-last tensor and y is dense channels-last of
    the same shape and dtype (0 otherwise)."""
    if g.dim() != 4 or g.dtype != y.dtype or g.shape != y.shape or not _is_cl(y) or g.is_contiguous(memory_format=_CL):
        return 0
    n, c, h, w = g.shape
    sn, sc, sh, sw = g.stride()
    if sc == 1 and sw > c and sh == w * sw and (n == 1 or sn == h * sh):
        return sw
    return 0


class _AddReLU(Function):
    """relu(a + b) as one launch (the tail of the encoders' residual blocks); the backward hands ONE masked gradient to both."""

    @staticmethod
    def forward(ctx, a, b):
        lib = _lib.load()
        out = torch.empty_like(a)
        _lib.check(lib.pcfa_add_relu_forward(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), a.numel(), 0 if a.dtype == torch.float32 else 1,
                                             _lib.stream()), "pcfa_add_relu_forward")
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (out,) = ctx.saved_tensors
        g = g.contiguous(memory_format=_CL) if _is_cl(out) else g.contiguous()
        gx = torch.empty_like(out)
        _lib.check(lib.pcfa_relu_mask_backward(_lib.ptr(out), _lib.ptr(g), _lib.ptr(gx), out.numel(), 0.0, 0 if out.dtype == torch.float32 else 1,
                                               _lib.stream()), "pcfa_relu_mask_backward")
        return gx, gx


class _AddReLUTwin(Function):
    """relu(a + b) returned TWICE (two views of one tensor) for a residual tail whose result feeds both the next block's
    convolution and its skip branch: the backward receives the two consumers' gradients separately and applies the ReLU mask
    to their sum in one pass (autograd would first add them: one more read-read-write pass over up to 58 MB)."""

    @staticmethod
    def forward(ctx, a, b):
        lib = _lib.load()
        out = torch.empty_like(a)
        _lib.check(lib.pcfa_add_relu_forward(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), a.numel(), 0, _lib.stream()), "pcfa_add_relu_forward")
        ctx.save_for_backward(out)
        ctx.set_materialize_grads(False)
        return out, out.view_as(out)

    @staticmethod
    def backward(ctx, g1, g2):
        lib = _lib.load()
        (out,) = ctx.saved_tensors
        if g1 is None and g2 is None:
            return None, None
        fmt = dict(memory_format=_CL) if _is_cl(out) else {}
        gx = torch.empty_like(out)
        if g1 is None or g2 is None:
            g = (g2 if g1 is None else g1).contiguous(**fmt)
            _lib.check(lib.pcfa_relu_mask_backward(_lib.ptr(out), _lib.ptr(g), _lib.ptr(gx), out.numel(), 0.0, 0, _lib.stream()),
                       "pcfa_relu_mask_backward")
        else:
            g1, g2 = g1.contiguous(**fmt), g2.contiguous(**fmt)
            _lib.check(lib.pcfa_relu_mask2_backward(_lib.ptr(out), _lib.ptr(g1), _lib.ptr(g2), _lib.ptr(gx), out.numel(), _lib.stream()),
                       "pcfa_relu_mask2_backward")
        return gx, gx


def add_relu(a: torch.Tensor, b: torch.Tensor, twin: bool = False) -> torch.Tensor:
    """relu(a + b); one kernel when both are CUDA tensors of the same dense layout (fp32 / fp16), torch ops otherwise.
    twin=True (fp32): the result carries a second autograd handle of itself in `._pcfa_twin`; a consumer that uses the value
    twice (a residual block: convolution + skip) takes one handle each, see _AddReLUTwin."""
    if (_ENABLED and a.is_cuda and a.shape == b.shape and a.dtype == b.dtype and a.dtype in (torch.float32, torch.float16)
            and a.stride() == b.stride() and (a.is_contiguous() or _is_cl(a)) and a.numel() % 8 == 0
            and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0):
        if twin and a.dtype == torch.float32 and torch.is_grad_enabled() and (a.requires_grad or b.requires_grad) \
                and os.environ.get("PCFA_ADD_RELU_TWIN", "1") != "0":
            out, out2 = _AddReLUTwin.apply(a, b)
            out._pcfa_twin = out2
            return out
        return _AddReLU.apply(a, b)
    return torch.relu(a + b)


class _FlowStep(Function):
    """coords1 + delta_flow and the next iteration's channels-last flow in one launch (csrc/bias_act.cu: flow_step_kernel)."""

    @staticmethod
    def forward(ctx, coords1, coords0, delta, flow_channels):
        lib = _lib.load()
        B, _, H, W = coords1.shape
        new = torch.empty_like(coords1)
        flow = torch.empty((B, flow_channels, H, W), device=coords1.device, dtype=torch.float32, memory_format=_CL)
        _lib.check(lib.pcfa_flow_step(_lib.ptr(coords1), _lib.ptr(coords0), _lib.ptr(delta), delta.stride(3),
                                      0 if delta.dtype == torch.float32 else 1, _lib.ptr(new), _lib.ptr(flow),
                                      flow_channels, B, H, W, _lib.stream()), "pcfa_flow_step")
        ctx.ddtype = delta.dtype
        ctx.mark_non_differentiable(flow)
        return new, flow

    @staticmethod
    def backward(ctx, gnew, gflow):
        if gnew is not None and gnew.dtype != ctx.ddtype:
            gnew = gnew.to(ctx.ddtype)
        return None, None, gnew, None                # d new_coords1 / d delta = identity (coords1 is detached by the caller)


def padded_flow(flow: torch.Tensor, channels: int) -> torch.Tensor:
    """[B, channels, H, W] channels-last with the flow in channels 0..1 and zeros behind (flow_step's output format)."""
    out = torch.zeros((flow.shape[0], channels, flow.shape[2], flow.shape[3]), device=flow.device, dtype=flow.dtype).contiguous(memory_format=_CL)
    out[:, :flow.shape[1]] = flow
    return out


def flow_step(coords1: torch.Tensor, coords0: torch.Tensor, delta: torch.Tensor, flow_channels: int = 2):
    """(coords1 + delta, channels-last (coords1 + delta - coords0) zero-padded to flow_channels).  delta: [B,2,H,W] view of
    a channels-last tensor (possibly a channel slice of the flow head's padded output)."""
    ok = (coords1.is_cuda and coords1.dtype == torch.float32 and delta.dtype in (torch.float32, torch.float16) and coords1.is_contiguous()
          and coords0.is_contiguous() and delta.dim() == 4 and delta.shape == coords1.shape and delta.stride(1) == 1
          and delta.stride(3) % 2 == 0 and delta.stride(2) == delta.shape[3] * delta.stride(3)
          and (delta.shape[0] == 1 or delta.stride(0) == delta.shape[2] * delta.stride(2)) and delta.data_ptr() % 8 == 0)
    if not (ok and _ENABLED):
        new = coords1 + delta.float().contiguous()
        return new, padded_flow((new - coords0).detach(), flow_channels)
    return _FlowStep.apply(coords1, coords0, delta, flow_channels)


class _ConvBiasAct(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, dilation, groups, relu, slope=0.0, tail=None, transposed=None):
        lib = _lib.load()
        if _CUDNN_FUSED and transposed is None and relu and slope == 0.0:
            # cuDNN's own convolution + bias + ReLU epilogue (one library call; 0.25 ms of the RAFT closure against the
            # convolution followed by the in-place epilogue kernel below)
            y = torch.ops.aten.cudnn_convolution_relu(x, weight, bias, stride, padding, dilation, groups)
            cl = _is_cl(y)
        else:
            if transposed is not None:               # ConvTranspose2d: `transposed` = its output_padding
                y = F.conv_transpose2d(x, weight, None, stride, padding, transposed, groups, dilation)
            else:
                y = F.conv2d(x, weight, None, stride, padding, dilation, groups)
            cl = _is_cl(y)
            if not (cl or y.is_contiguous()):
                y = y.contiguous()
            C = y.shape[1]
            inner = 1 if cl else y.shape[2] * y.shape[3]
            st = lib.pcfa_bias_act_forward(_lib.ptr(y), _lib.ptr(bias), y.numel(), C, inner, int(relu), float(slope),
                                           0 if y.dtype == torch.float32 else 1, _lib.stream())
            if st == -1:                             # PCFA_E_BADARG: shape the vector kernels do not take (e.g. 2 output channels)
                y.add_(bias.view(1, -1, 1, 1))
                if relu:
                    y = F.leaky_relu_(y, slope) if slope else y.relu_()
            else:
                _lib.check(st, "pcfa_bias_act_forward")
        if tail is not None:                         # overwrite the last channels (zero filters there): a cat without the copy
            y[:, y.shape[1] - tail.shape[1]:] = tail
        ctx.save_for_backward(weight, y if relu else None)
        ctx.meta = (tuple(x.shape), x.dtype, x.device, _is_cl(x) or cl, stride, padding, dilation, groups, bool(relu), float(slope), transposed)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        weight, y = ctx.saved_tensors
        shape, dtype, device, cl, stride, padding, dilation, groups, relu, slope, transposed = ctx.meta
        if relu:
            ld = _channel_slice_ld(g, y)
            gx = torch.empty_like(y)
            if ld:                                   # g is a channel slice of a wider NHWC tensor (gradient of a concatenation)
                st = lib.pcfa_relu_mask_backward_rows(_lib.ptr(y), _lib.ptr(g), _lib.ptr(gx), y.numel() // y.shape[1], y.shape[1], ld,
                                                      slope, 0 if y.dtype == torch.float32 else 1, _lib.stream())
            else:
                st = -1
            if st == -1:
                g = g.contiguous(memory_format=_CL) if _is_cl(y) else g.contiguous()
                if g.dtype != y.dtype:
                    g = g.to(y.dtype)
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
                                                  transposed is not None, (0, 0) if transposed is None else transposed, groups,
                                                  (True, False, False))[0]
        return gin, None, None, None, None, None, None, None, None, None, None


class _Fork(Function):
    """x -> (x, x) for a tensor with two consumers, one of which is a channels-last concatenation: the backward receives the
    two gradients SEPARATELY (autograd would sum a dense tensor and a strided channel slice with ATen's non-vectorised add)
    and adds them with one vectorised kernel."""

    @staticmethod
    def forward(ctx, x):
        ctx.set_materialize_grads(False)
        return x.view_as(x), x.view_as(x)

    @staticmethod
    def backward(ctx, g1, g2):
        if g1 is None or g2 is None:
            return g2 if g1 is None else g1
        for a, b in ((g1, g2), (g2, g1)):
            if a.dtype == torch.float32 and _is_cl(a) and a.shape[1] % 4 == 0:
                ld = _channel_slice_ld(b, a)
                if ld and ld % 4 == 0 and b.data_ptr() % 16 == 0:
                    out = torch.empty_like(a)
                    _lib.check(_lib.load().pcfa_add_rows(_lib.ptr(out), _lib.ptr(a), _lib.ptr(b), a.numel() // a.shape[1], a.shape[1], ld,
                                                         _lib.stream()), "pcfa_add_rows")
                    return out
        return g1 + g2


def fork(x: torch.Tensor):
    """(x, x) — give one to the next convolution and the other to the concatenation that follows later (see _Fork)."""
    if _ENABLED and x.is_cuda and x.requires_grad and x.dtype == torch.float32 and x.dim() == 4 and os.environ.get("PCFA_FORK", "1") != "0":
        return _Fork.apply(x)
    return x, x


class _DenseConvCat(Function):
    """x -> cat(act(conv(x) + b), x) along the channels (channels-last, stride-1 convolution with frozen weights) as ONE
    autograd node: PWCNet's DenseNet decoder (PWCNet.py:253-257).  Backward: the activation mask reads its gradient straight
    from the first channels of the concatenation's gradient, cuDNN's data gradient follows, and the skip branch's share — the
    remaining channels, a strided slice — is added in place by one vectorised kernel.  autograd otherwise sums a strided and a
    dense tensor with ATen's non-vectorised add (51 launches, 0.34 ms of a 4.0 ms PWCNet closure)."""

    @staticmethod
    def forward(ctx, x, weight, bias, padding, slope):
        import ctypes as C
        lib = _lib.load()
        y = F.conv2d(x, weight, None, 1, padding)
        _lib.check(lib.pcfa_bias_act_forward(_lib.ptr(y), _lib.ptr(bias), y.numel(), y.shape[1], 1, 1, float(slope), 0, _lib.stream()),
                   "pcfa_bias_act_forward")
        B, cy, H, W = y.shape
        cx = x.shape[1]
        out = torch.empty((B, cy + cx, H, W), device=x.device, dtype=torch.float32, memory_format=_CL)
        ptrs = (C.c_void_p * 2)(y.data_ptr(), x.data_ptr())
        chans = (C.c_int * 2)(cy, cx)
        _lib.hint_bytes(2 * 4 * out.numel())
        _lib.check(lib.pcfa_cat_channels_last(C.cast(ptrs, C.c_void_p), C.cast(chans, C.c_void_p), 2, _lib.ptr(out), B * H * W, _lib.stream()),
                   "pcfa_cat_channels_last")
        ctx.save_for_backward(weight, y)
        ctx.meta = (tuple(x.shape), padding, float(slope))
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        weight, y = ctx.saved_tensors
        xshape, padding, slope = ctx.meta
        g = g.contiguous(memory_format=_CL)
        B, cy, H, W = y.shape
        ctot, cx = g.shape[1], xshape[1]
        gy = torch.empty_like(y)
        _lib.check(lib.pcfa_relu_mask_backward_rows(_lib.ptr(y), _lib.ptr(g), _lib.ptr(gy), B * H * W, cy, ctot, slope, 0, _lib.stream()),
                   "pcfa_relu_mask_backward_rows")
        gx = torch.ops.aten.convolution_backward(gy, _dummy_like(xshape, y.dtype, y.device, True), weight, None, (1, 1), padding, (1, 1),
                                                 False, (0, 0), 1, (True, False, False))[0]
        gx = gx.contiguous(memory_format=_CL)
        _lib.check(lib.pcfa_add_rows_inplace(_lib.ptr(gx), C_void(g.data_ptr() + 4 * cy), B * H * W, cx, ctot, _lib.stream()),
                   "pcfa_add_rows_inplace")
        return gx, None, None, None, None


def C_void(addr: int):
    import ctypes
    return ctypes.c_void_p(addr)


def dense_conv_cat(conv: torch.nn.Conv2d, x: torch.Tensor, slope: float):
    """torch.cat((LeakyReLU(conv(x)), x), 1) — one autograd node on the channels-last GPU path, torch ops otherwise."""
    frozen = not (conv.weight.requires_grad or (conv.bias is not None and conv.bias.requires_grad))
    if (_ENABLED and x.is_cuda and x.dtype == torch.float32 and _is_cl(x) and frozen and conv.bias is not None
            and type(conv) is torch.nn.Conv2d and conv.stride == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1
            and conv.padding_mode == "zeros" and x.shape[1] % 4 == 0 and conv.out_channels % 4 == 0
            and os.environ.get("PCFA_DENSE_CAT", "1") != "0"):
        w = padded_in_channels(conv, x.shape[1]) if x.shape[1] != conv.in_channels else conv.weight
        if _is_cl(w) or w.shape[2:] == (1, 1):
            return _DenseConvCat.apply(x, w, conv.bias, tuple(conv.padding), float(slope))
    y = apply_conv(conv, x, slope)
    return torch.cat((y, x), 1)


def padded_out_channels(conv: torch.nn.Conv2d, multiple: int = 8):
    """(weight, bias) of a frozen convolution with zero filters appended so that the output channel count is a multiple of
    `multiple`: cuDNN's sm_100 NHWC kernels otherwise wrap the convolution (and its data gradient) in channel-padding
    launches (nhwcAddPaddingKernel: 76 launches, 0.28 ms per RAFT closure for the 126- and 2-channel outputs).  Cached."""
    key = (conv.weight.data_ptr(), conv.weight._version, conv.bias.data_ptr(), conv.bias._version, multiple)
    cache = getattr(conv, "_pcfa_padded", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            co = conv.weight.shape[0]
            extra = (-co) % multiple
            w = torch.cat([conv.weight, conv.weight.new_zeros((extra,) + tuple(conv.weight.shape[1:]))], 0)
            if _is_cl(conv.weight):
                w = w.contiguous(memory_format=_CL)
            b = torch.cat([conv.bias, conv.bias.new_zeros(extra)], 0).contiguous()
        cache = (key, w, b)
        conv._pcfa_padded = cache
    return cache[1], cache[2]


def padded_in_channels(conv, cin_padded: int):
    """Weight of a frozen Conv2d / ConvTranspose2d with zero INPUT channels appended up to cin_padded (its input is a
    concatenation that cat_channels(pad_to=...) extended with zero channels).  Cached per (module, cin_padded)."""
    key = (conv.weight.data_ptr(), conv.weight._version, cin_padded)
    cache = getattr(conv, "_pcfa_padded_in", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            w = conv.weight
            dim = 0 if isinstance(conv, torch.nn.ConvTranspose2d) else 1
            extra = cin_padded - w.shape[dim]
            assert extra >= 0
            shape = list(w.shape)
            shape[dim] = extra
            wp = torch.cat([w, w.new_zeros(shape)], dim)
            wp = wp.contiguous(memory_format=_CL) if _is_cl(w) else wp.contiguous()
        cache = (key, wp)
        conv._pcfa_padded_in = cache
    return cache[1]


def conv_act(conv: torch.nn.Conv2d, x: torch.Tensor, relu: bool, weight=None, bias=None, tag: str = "_pcfa_w16", slope: float = 0.0,
             tail=None):
    """act?(conv(x)) with `conv`'s geometry (nn.Conv2d or nn.ConvTranspose2d) (act = ReLU, or LeakyReLU(slope) for slope > 0); `weight` / `bias` override the
    module's (e.g. batch-norm-folded copies).  `tail` ([B, k, H, W], no gradient): written over the LAST k output channels
    (which the caller has given zero filters, see padded_out_channels) — torch.cat([conv_out, tail], 1) without the copy."""
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
            tr = tuple(conv.output_padding) if isinstance(conv, torch.nn.ConvTranspose2d) else None
            return _ConvBiasAct.apply(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups, relu, slope, tail, tr)
    if isinstance(conv, torch.nn.ConvTranspose2d):
        y = F.conv_transpose2d(x, w, b, conv.stride, conv.padding, conv.output_padding, conv.groups, conv.dilation)
    else:
        y = F.conv2d(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
    y = (F.leaky_relu(y, slope) if slope else F.relu(y)) if relu else y
    if tail is not None:
        y = torch.cat([y[:, :y.shape[1] - tail.shape[1]], tail], 1)
    return y


def apply_conv(conv, x, act_slope=None):
    """A (frozen) Conv2d / ConvTranspose2d on an input that cat_channels(pad_to=...) may have extended with zero channels:
    zero-padded input-channel weights, bias (and LeakyReLU(act_slope) / ReLU for act_slope == 0) as the fused epilogue."""
    w = padded_in_channels(conv, x.shape[1]) if x.shape[1] != conv.in_channels else None
    if conv.bias is None:
        if isinstance(conv, torch.nn.ConvTranspose2d):
            y = F.conv_transpose2d(x, conv.weight if w is None else w, None, conv.stride, conv.padding, conv.output_padding, conv.groups, conv.dilation)
        else:
            y = F.conv2d(x, conv.weight if w is None else w, None, conv.stride, conv.padding, conv.dilation, conv.groups)
        return y if act_slope is None else (F.leaky_relu(y, act_slope) if act_slope else F.relu(y))
    return conv_act(conv, x, act_slope is not None, weight=w, tag="_pcfa_w16" if w is None else "_pcfa_padin16", slope=act_slope or 0.0)


class ConvLeakyReLU(torch.nn.Sequential):
    """nn.Sequential(Conv2d | ConvTranspose2d, LeakyReLU(slope)) with the same children and state-dict keys (PWCNet.py:23-28,
    FlowNet's submodules.py conv() / deconv()), evaluated through conv_act when the weights are frozen and the input is a
    CUDA tensor."""

    def forward(self, x):
        return apply_conv(self[0], x, float(self[1].negative_slope))


class ConvOnly(torch.nn.Sequential):
    """nn.Sequential(Conv2d) (FlowNet's i_conv) on possibly zero-padded inputs."""

    def forward(self, x):
        return apply_conv(self[0], x, None)
