"""Fused element-wise halves of RAFT/GMA's convolutional GRU (models/raft/update.py:16-60) — row f-4 of SURVEY.md
section 8.  `gru_gates(zr, h) -> (z, rh)` and `gru_blend(z, q_pre, h) -> h_new` replace eight ATen launches per GRU step
forward and about ten backward with two each; CUDA fp32 NCHW-contiguous tensors only."""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import _lib


def _is_cl(t: torch.Tensor) -> bool:
    return t.dim() == 4 and not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last)


def _check(*ts, name):
    for t in ts:
        if not t.is_cuda or t.dtype != torch.float32:
            raise RuntimeError(f"{name}: expected CUDA float32 tensors (pcfa_b200 has no CPU path)")


class _Gates(Function):
    @staticmethod
    def forward(ctx, zr, h):
        cl = _is_cl(h)
        fmt = torch.channels_last if cl else torch.contiguous_format
        zr, h = zr.contiguous(memory_format=fmt), h.contiguous(memory_format=fmt)
        _check(zr, h, name="gru_gates")
        B, C = h.shape[0], h.shape[1]
        if zr.shape[1] != 2 * C or zr.shape[0] != B or zr.shape[2:] != h.shape[2:]:
            raise RuntimeError(f"gru_gates: zr {tuple(zr.shape)} does not match h {tuple(h.shape)}")
        n = h.numel() // B
        z, r, rh = torch.empty_like(h), torch.empty_like(h), torch.empty_like(h)
        lib = _lib.load()
        _lib.check(lib.pcfa_gru_gates_forward(_lib.ptr(zr), _lib.ptr(h), _lib.ptr(z), _lib.ptr(r), _lib.ptr(rh), B, n,
                                              C if cl else 0, _lib.stream()), "pcfa_gru_gates_forward")
        ctx.save_for_backward(z, r, h)
        ctx.cl = cl
        return z, rh

    @staticmethod
    def backward(ctx, gz, grh):
        z, r, h = ctx.saved_tensors
        B = h.shape[0]
        n = h.numel() // B
        fmt = torch.channels_last if ctx.cl else torch.contiguous_format
        gz = None if gz is None else gz.contiguous(memory_format=fmt)
        grh = None if grh is None else grh.contiguous(memory_format=fmt)
        gzr = torch.empty((B, 2 * h.shape[1], *h.shape[2:]), device=h.device, dtype=h.dtype, memory_format=fmt)
        gh = torch.empty_like(h)
        lib = _lib.load()
        _lib.check(lib.pcfa_gru_gates_backward(_lib.ptr(z), _lib.ptr(r), _lib.ptr(h), _lib.ptr(gz), _lib.ptr(grh),
                                               _lib.ptr(gzr), _lib.ptr(gh), B, n, h.shape[1] if ctx.cl else 0,
                                               _lib.stream()), "pcfa_gru_gates_backward")
        return gzr, gh


class _Blend(Function):
    @staticmethod
    def forward(ctx, z, q_pre, h):
        fmt = torch.channels_last if _is_cl(h) else torch.contiguous_format      # element-wise: any common dense layout
        z, q_pre, h = z.contiguous(memory_format=fmt), q_pre.contiguous(memory_format=fmt), h.contiguous(memory_format=fmt)
        _check(z, q_pre, h, name="gru_blend")
        ctx.fmt = fmt
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
        ghn = ghn.contiguous(memory_format=ctx.fmt)
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


class _CatCL(Function):
    @staticmethod
    def forward(ctx, *xs):
        import ctypes as C
        xs = [x.contiguous(memory_format=torch.channels_last) for x in xs]
        _check(*xs, name="cat_channels_last")
        B, _, H, W = xs[0].shape
        cs = [int(x.shape[1]) for x in xs]
        out = torch.empty((B, sum(cs), H, W), device=xs[0].device, dtype=torch.float32, memory_format=torch.channels_last)
        ptrs = (C.c_void_p * len(xs))(*[x.data_ptr() for x in xs])
        chans = (C.c_int * len(xs))(*cs)
        lib = _lib.load()
        _lib.hint_bytes(2 * 4 * out.numel())
        _lib.check(lib.pcfa_cat_channels_last(C.cast(ptrs, C.c_void_p), C.cast(chans, C.c_void_p), len(xs), _lib.ptr(out),
                                              B * H * W, _lib.stream()), "pcfa_cat_channels_last")
        ctx.cs = cs
        return out

    @staticmethod
    def backward(ctx, g):
        return tuple(torch.split(g, ctx.cs, dim=1))


def cat_channels(xs, channels_last: bool):
    """torch.cat(xs, dim=1); channels-last inputs (B x C_k x H x W, at most four) go through one vectorised kernel."""
    if channels_last and len(xs) <= 4 and all(x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 for x in xs):
        return _CatCL.apply(*xs)
    return torch.cat(xs, dim=1)


_CL = torch.channels_last


class _GatesX(Function):
    """z, [r*h | m] from zr + addend (channels-last; see pcfa_gru_gates_x_forward)."""

    @staticmethod
    def forward(ctx, zr, addend, h, m):
        zr, addend, h, m = (t.contiguous(memory_format=_CL) for t in (zr, addend, h, m))
        _check(zr, addend, h, m, name="gru_gates_x")
        B, C, H, W = h.shape
        Cm = m.shape[1]
        z, r = torch.empty_like(h), torch.empty_like(h)
        rhm = torch.empty((B, C + Cm, H, W), device=h.device, dtype=h.dtype, memory_format=_CL)
        lib = _lib.load()
        _lib.check(lib.pcfa_gru_gates_x_forward(_lib.ptr(zr), _lib.ptr(addend), _lib.ptr(h), _lib.ptr(m), _lib.ptr(z), _lib.ptr(r),
                                                _lib.ptr(rhm), C, Cm, B * H * W, _lib.stream()), "pcfa_gru_gates_x_forward")
        ctx.save_for_backward(z, r, h)
        ctx.cm = Cm
        return z, rhm

    @staticmethod
    def backward(ctx, gz, grhm):
        z, r, h = ctx.saved_tensors
        B, C, H, W = h.shape
        gz = None if gz is None else gz.contiguous(memory_format=_CL)
        grhm = None if grhm is None else grhm.contiguous(memory_format=_CL)
        gzr = torch.empty((B, 2 * C, H, W), device=h.device, dtype=h.dtype, memory_format=_CL)
        gh = torch.empty_like(h)
        lib = _lib.load()
        _lib.check(lib.pcfa_gru_gates_x_backward(_lib.ptr(z), _lib.ptr(r), _lib.ptr(h), _lib.ptr(gz), _lib.ptr(grhm), _lib.ptr(gzr),
                                                 _lib.ptr(gh), C, ctx.cm, B * H * W, _lib.stream()), "pcfa_gru_gates_x_backward")
        return gzr, gzr, gh, (None if grhm is None else grhm[:, C:])


class _BlendX(Function):
    """h_new (and [h_new | m]) from z, q_pre + addend, h (channels-last; see pcfa_gru_blend_x_forward)."""

    @staticmethod
    def forward(ctx, z, q_pre, addend, h, m, make_hm):
        z, q_pre, addend, h, m = (t.contiguous(memory_format=_CL) for t in (z, q_pre, addend, h, m))
        _check(z, q_pre, addend, h, m, name="gru_blend_x")
        B, C, H, W = h.shape
        Cm = m.shape[1]
        q, hn = torch.empty_like(h), torch.empty_like(h)
        hm = torch.empty((B, C + Cm, H, W), device=h.device, dtype=h.dtype, memory_format=_CL) if make_hm else None
        lib = _lib.load()
        _lib.check(lib.pcfa_gru_blend_x_forward(_lib.ptr(z), _lib.ptr(q_pre), _lib.ptr(addend), _lib.ptr(h), _lib.ptr(m), _lib.ptr(q),
                                                _lib.ptr(hn), _lib.ptr(hm), C, Cm, B * H * W, _lib.stream()), "pcfa_gru_blend_x_forward")
        ctx.save_for_backward(z, q, h)
        ctx.cm, ctx.make_hm = Cm, bool(make_hm)
        return (hn, hm) if make_hm else hn

    @staticmethod
    def backward(ctx, ghn, ghm=None):
        z, q, h = ctx.saved_tensors
        B, C, H, W = h.shape
        ghn = None if ghn is None else ghn.contiguous(memory_format=_CL)
        ghm = None if ghm is None else ghm.contiguous(memory_format=_CL)
        gz, gq, gh = torch.empty_like(h), torch.empty_like(h), torch.empty_like(h)
        lib = _lib.load()
        _lib.check(lib.pcfa_gru_blend_x_backward(_lib.ptr(z), _lib.ptr(q), _lib.ptr(h), _lib.ptr(ghn), _lib.ptr(ghm), _lib.ptr(gz),
                                                 _lib.ptr(gq), _lib.ptr(gh), C, ctx.cm, B * H * W, _lib.stream()), "pcfa_gru_blend_x_backward")
        return gz, gq, gq, gh, (None if ghm is None else ghm[:, C:]), None


def gru_gates_x(zr, addend, h, m):
    return _GatesX.apply(zr, addend, h, m)


def gru_blend_x(z, q_pre, addend, h, m, make_hm: bool):
    return _BlendX.apply(z, q_pre, addend, h, m, make_hm)
