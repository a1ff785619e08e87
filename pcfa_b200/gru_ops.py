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
    def forward(ctx, pad_to, *xs):
        import ctypes as C
        xs = [x.contiguous(memory_format=torch.channels_last) for x in xs]
        _check(*xs, name="cat_channels_last")
        B, _, H, W = xs[0].shape
        cs = [int(x.shape[1]) for x in xs]
        total = sum(cs)
        ctot = total if not pad_to else -(-total // pad_to) * pad_to
        out = torch.empty((B, ctot, H, W), device=xs[0].device, dtype=torch.float32, memory_format=torch.channels_last)
        ptrs = (C.c_void_p * len(xs))(*[x.data_ptr() for x in xs])
        chans = (C.c_int * len(xs))(*cs)
        lib = _lib.load()
        _lib.hint_bytes(2 * 4 * out.numel())
        if pad_to:
            _lib.check(lib.pcfa_cat_channels_last_pad(C.cast(ptrs, C.c_void_p), C.cast(chans, C.c_void_p), len(xs), _lib.ptr(out),
                                                      B * H * W, ctot, _lib.stream()), "pcfa_cat_channels_last_pad")
        else:
            _lib.check(lib.pcfa_cat_channels_last(C.cast(ptrs, C.c_void_p), C.cast(chans, C.c_void_p), len(xs), _lib.ptr(out),
                                                  B * H * W, _lib.stream()), "pcfa_cat_channels_last")
        ctx.cs = cs
        return out

    @staticmethod
    def backward(ctx, g):
        return (None,) + tuple(torch.split(g[:, :sum(ctx.cs)] if g.shape[1] != sum(ctx.cs) else g, ctx.cs, dim=1))


def cat_channels(xs, channels_last: bool, pad_to: int = 0):
    """torch.cat(xs, dim=1); channels-last inputs (B x C_k x H x W, at most four) go through one vectorised kernel.
    pad_to > 0 (a multiple of 4): zero channels are appended so that the channel count is a multiple of pad_to (the consumer
    convolution gets zero-padded input-channel weights, conv_ops.padded_in_channels)."""
    if channels_last and len(xs) <= 4 and all(x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 for x in xs):
        return _CatCL.apply(int(pad_to), *xs)
    y = torch.cat(xs, dim=1)
    if pad_to and y.shape[1] % pad_to:
        y = torch.nn.functional.pad(y, (0, 0, 0, 0, 0, (-y.shape[1]) % pad_to))
    return y


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


# ------------------------------------------------------------------------------------------------------------------
# One autograd node per SepConvGRU step (horizontal + vertical half step) on the NHWC / hoisted path.
class _HoistSource(Function):
    """Identity on a hoisted addend P (conv(inp, W_inp) + bias, computed once per forward).  Every GRU step consumes the
    returned tensor, returns no gradient for it, and instead accumulates its share into `acc` inside its own backward
    kernels (acc = on the last iteration, acc += on the others: the last iteration's backward runs first).  This node is
    scheduled by autograd after all of them and hands `acc` on — the sum over iterations without 11 ATen adds per addend."""

    @staticmethod
    def forward(ctx, P, acc):
        ctx.acc = acc
        ctx.set_materialize_grads(False)
        return P.view_as(P)

    @staticmethod
    def backward(ctx, g):
        acc = ctx.acc
        ctx.acc = None
        return (acc if g is None else acc + g), None


def hoist_sources(hoist):
    """[(W_zr, P_zr, W_q, P_q, pad)] x 2 -> the same with P wrapped by _HoistSource, plus the accumulators.  Under fp16
    autocast (GMA) the addends are half: the frozen weights are cast once here instead of by every convolution call."""
    out = []
    for (wzr, pzr, wq, pq, pad) in hoist:
        pzr, pq = pzr.contiguous(memory_format=_CL), pq.contiguous(memory_format=_CL)
        if pzr.dtype == torch.float16:
            wzr = wzr.to(torch.float16).contiguous(memory_format=_CL)
            wq = wq.to(torch.float16).contiguous(memory_format=_CL)
        azr, aq = torch.empty_like(pzr), torch.empty_like(pq)
        out.append((wzr, _HoistSource.apply(pzr, azr), wq, _HoistSource.apply(pq, aq), pad, azr, aq))
    return out


_DGRAD_DUMMY = {}


def _dgrad(gout, weight, in_channels, pad):
    """Data gradient of a stride-1 convolution with frozen weights (channels-last): cuDNN dgrad, nothing else."""
    B, _, H, W = gout.shape
    key = (gout.device, gout.dtype, B, in_channels, H, W)
    dummy = _DGRAD_DUMMY.get(key)
    if dummy is None:                                  # only its sizes / memory format are read
        dummy = torch.empty((B, in_channels, H, W), device=gout.device, dtype=gout.dtype, memory_format=_CL)
        _DGRAD_DUMMY[key] = dummy
    return torch.ops.aten.convolution_backward(gout, dummy, weight, None, (1, 1), pad, (1, 1), False, (0, 0), 1,
                                               (True, False, False))[0]


class _GRUStepX(Function):
    @staticmethod
    def forward(ctx, h, m, pzr1, pq1, pzr2, pq2, wzr1, wq1, wzr2, wq2, pad1, pad2, accs, acc_mode):
        import ctypes as C
        import torch.nn.functional as F
        lib, P, s = _lib.load(), _lib.ptr, _lib.stream()
        dt = pzr1.dtype                                   # fp32 (RAFT) or fp16 (GMA under autocast): one dtype throughout
        if dt not in (torch.float32, torch.float16) or not h.is_cuda:
            raise RuntimeError("gru_step_x: expected CUDA float32 / float16 tensors (pcfa_b200 has no CPU path)")
        sfx = "_h" if dt == torch.float16 else ""
        h, m = h.to(dt).contiguous(memory_format=_CL), m.to(dt).contiguous(memory_format=_CL)
        B, Ch, H, W = h.shape
        Cm = m.shape[1]
        npix = B * H * W

        def new(c):
            return torch.empty((B, c, H, W), device=h.device, dtype=dt, memory_format=_CL)
        hm = new(Ch + Cm)
        if sfx:
            _lib.check(lib.pcfa_cat2_channels_last_h(P(h), P(m), P(hm), Ch, Cm, npix, s), "pcfa_cat2_channels_last_h")
        else:
            ptrs = (C.c_void_p * 2)(h.data_ptr(), m.data_ptr())
            chans = (C.c_int * 2)(Ch, Cm)
            _lib.check(lib.pcfa_cat_channels_last(C.cast(ptrs, C.c_void_p), C.cast(chans, C.c_void_p), 2, P(hm), npix, s), "pcfa_cat_channels_last")
        gates_fwd, blend_fwd = getattr(lib, "pcfa_gru_gates_x_forward" + sfx), getattr(lib, "pcfa_gru_blend_x_forward" + sfx)
        saved = []
        hin = h
        for (wzr, pzr, wq, pq, pad, last) in ((wzr1, pzr1, wq1, pq1, pad1, False), (wzr2, pzr2, wq2, pq2, pad2, True)):
            zr = F.conv2d(hm, wzr, None, 1, pad)
            z, r, rhm = new(Ch), new(Ch), new(Ch + Cm)
            _lib.check(gates_fwd(P(zr), P(pzr), P(hin), P(m), P(z), P(r), P(rhm), Ch, Cm, npix, s), "pcfa_gru_gates_x_forward" + sfx)
            qp = F.conv2d(rhm, wq, None, 1, pad)
            q, hn = new(Ch), new(Ch)
            hm = None if last else new(Ch + Cm)
            _lib.check(blend_fwd(P(z), P(qp), P(pq), P(hin), P(m), P(q), P(hn), P(hm), Ch, Cm, npix, s), "pcfa_gru_blend_x_forward" + sfx)
            saved += [z, r, q, hin]
            hin = hn
        ctx.save_for_backward(*saved, wzr1, wq1, wzr2, wq2)
        ctx.pads, ctx.accs, ctx.acc_mode, ctx.cm, ctx.sfx = (pad1, pad2), accs, int(acc_mode), Cm, sfx
        return hin

    @staticmethod
    def backward(ctx, gh2):
        lib, P, s = _lib.load(), _lib.ptr, _lib.stream()
        z1, r1, q1, h0, z2, r2, q2, h1, wzr1, wq1, wzr2, wq2 = ctx.saved_tensors
        (pad1, pad2), (azr1, aq1, azr2, aq2), mode, Cm = ctx.pads, ctx.accs, ctx.acc_mode, ctx.cm
        B, Ch, H, W = h0.shape
        npix = B * H * W
        sfx = ctx.sfx
        gh2 = gh2.to(h0.dtype).contiguous(memory_format=_CL)
        blend_bwd, gates_bwd = getattr(lib, "pcfa_gru_blend_x_backward_acc" + sfx), getattr(lib, "pcfa_gru_gates_x_backward_acc" + sfx)
        combine = getattr(lib, "pcfa_gru_step_combine" + sfx)

        def new(c):
            return torch.empty((B, c, H, W), device=h0.device, dtype=h0.dtype, memory_format=_CL)
        # ---- vertical half step (second)
        gz, gq, gh1_a = new(Ch), new(Ch), new(Ch)
        _lib.check(blend_bwd(P(z2), P(q2), P(h1), P(gh2), None, None, P(gz), P(gq), P(gh1_a), P(aq2), mode,
                                                     Ch, Cm, npix, s), "pcfa_gru_blend_x_backward_acc")
        grhm2 = _dgrad(gq, wq2, Ch + Cm, pad2)
        gzr, gh1_b = new(2 * Ch), new(Ch)
        _lib.check(gates_bwd(P(z2), P(r2), P(h1), P(gz), P(grhm2), P(gzr), P(gh1_b), P(azr2), mode, Ch, Cm, npix, s),
                   "pcfa_gru_gates_x_backward_acc")
        ghm2 = _dgrad(gzr, wzr2, Ch + Cm, pad2)
        # ---- horizontal half step (first): grad of h1 = gh1_a + gh1_b + ghm2[:, :C]
        gz, gq, gh0_a = new(Ch), new(Ch), new(Ch)
        _lib.check(blend_bwd(P(z1), P(q1), P(h0), P(gh1_a), P(gh1_b), P(ghm2), P(gz), P(gq), P(gh0_a), P(aq1), mode,
                                                     Ch, Cm, npix, s), "pcfa_gru_blend_x_backward_acc")
        grhm1 = _dgrad(gq, wq1, Ch + Cm, pad1)
        gzr, gh0_b = new(2 * Ch), new(Ch)
        _lib.check(gates_bwd(P(z1), P(r1), P(h0), P(gz), P(grhm1), P(gzr), P(gh0_b), P(azr1), mode, Ch, Cm, npix, s),
                   "pcfa_gru_gates_x_backward_acc")
        ghm1 = _dgrad(gzr, wzr1, Ch + Cm, pad1)
        gh, gm = new(Ch), new(Cm)
        _lib.check(combine(P(gh0_a), P(gh0_b), P(ghm1), P(grhm1), P(grhm2), P(ghm2), P(gh), P(gm), Ch, Cm, npix, s),
                   "pcfa_gru_step_combine")
        ctx.accs = None
        return (gh, gm) + (None,) * 12


def gru_step_x(h, m, sources, last_iteration: bool):
    """h' = SepConvGRU(h, [inp | m]) on the hoisted NHWC path as ONE autograd node; `sources` from hoist_sources()."""
    (wzr1, pzr1, wq1, pq1, pad1, azr1, aq1), (wzr2, pzr2, wq2, pq2, pad2, azr2, aq2) = sources
    return _GRUStepX.apply(h, m, pzr1, pq1, pzr2, pq2, wzr1, wq1, wzr2, wq2, tuple(pad1), tuple(pad2), (azr1, aq1, azr2, aq2),
                           1 if last_iteration else 2)
