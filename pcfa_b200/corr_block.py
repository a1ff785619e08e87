"""RAFT / GMA all-pairs correlation pyramid + multi-level lookup on B200.

Drop-in for `CorrBlock` of the reference (models/raft/corr.py:12-60, models/gma/corr.py:15-63):

    corr_fn = CorrBlock(fmap1, fmap2, num_levels=4, radius=4)
    corr    = corr_fn(coords)            # [B, num_levels*(2r+1)^2, H, W] float32 contiguous
    corr_fn.corr_pyramid                 # list of [B*H*W, 1, H_l, W_l]
    CorrBlock.corr(fmap1, fmap2)         # [B, H, W, 1, H, W]

Differences in mechanism (not in results): the pyramid is one flat buffer written by a single
build call; each lookup is one fused kernel over all levels; the backward of the lookups scatters
into ONE persistent gradient pyramid (zeroed once per backward pass) instead of autograd summing a
dense 261 MB gradient per lookup, and the build's backward contracts that gradient pyramid with
the (pooled) feature maps without folding it to level 0.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib


def pyramid_layout(B, H, W, num_levels):
    """(offsets[num_levels+1], hs, ws) of the flat pyramid buffer — same code path as the kernels."""
    lib = _lib.load()
    offs = (C.c_int64 * (num_levels + 1))()
    hs = (C.c_int * num_levels)()
    ws = (C.c_int * num_levels)()
    _lib.check(lib.pcfa_corr_pyramid_layout(B, H, W, num_levels, offs, hs, ws), "pcfa_corr_pyramid_layout")
    return list(offs), list(hs), list(ws)


def _impl_from_env():
    return int(os.environ.get("PCFA_CORR_IMPL", "0"))


class _PyramidState:
    """Per-CorrBlock side state shared by the build and lookup autograd nodes."""
    __slots__ = ("B", "C", "H", "W", "levels", "total", "grad", "impl", "occ_words", "lookups")

    def __init__(self, B, Cc, H, W, levels, total, impl):
        self.B, self.C, self.H, self.W, self.levels, self.total, self.impl = B, Cc, H, W, levels, total, impl
        self.grad = None        # persistent gradient pyramid of the current backward pass, + the occupancy bitmap as its tail
        self.occ_words = 0      # 32-bit words of the bitmap behind the pyramid in `grad` (0: none)
        self.lookups = []       # (coords, radius) of every lookup that has accumulated into `grad`

    def new_grad(self, device):
        """Zeroed gradient pyramid with the occupancy bitmap of the sparse backward behind it: ONE fill clears both."""
        lib = _lib.load()
        sparse = os.environ.get("PCFA_BWD_SPARSE", "1") != "0"
        self.occ_words = (lib.pcfa_corr_occupancy_bytes(self.B, self.H, self.W, self.levels) // 4) if sparse else 0
        self.lookups = []
        self.grad = torch.zeros(self.total + self.occ_words, device=device, dtype=torch.float32)


class _BuildPyramid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fmap1, fmap2, state):
        lib = _lib.load()
        _lib.require_cuda(fmap1, fmap2, name="CorrBlock")
        B, Cc, H, W, L = state.B, state.C, state.H, state.W, state.levels
        pyr = torch.empty(state.total, device=fmap1.device, dtype=torch.float32)
        wsb = lib.pcfa_corr_pyramid_workspace_bytes(B, Cc, H, W, L)
        ws = torch.empty(wsb, device=fmap1.device, dtype=torch.uint8)
        _lib.check(lib.pcfa_corr_pyramid_forward(_lib.ptr(fmap1), _lib.ptr(fmap2), _lib.ptr(pyr),
                                                 _lib.ptr(ws), wsb, B, Cc, H, W, L, state.impl,
                                                 _lib.stream()), "pcfa_corr_pyramid_forward")
        ctx.save_for_backward(fmap1, fmap2)
        ctx.state = state
        ctx.set_materialize_grads(False)
        return pyr

    @staticmethod
    def backward(ctx, gpyr):
        lib = _lib.load()
        fmap1, fmap2 = ctx.saved_tensors
        st = ctx.state
        G, lookups = st.grad, st.lookups
        st.grad, st.lookups = None, []
        if G is None and gpyr is None:
            return None, None, None
        occ = None
        if G is None:
            G = gpyr.contiguous()
        elif gpyr is not None:          # someone differentiated through corr_pyramid directly: dense gradient
            G = G[:st.total] + gpyr
        elif st.occ_words:
            # only the lookups wrote into G: mark the 32x32 blocks their windows cover (one launch per radius in use) and let
            # the build's backward skip the rest (pcfa_b200.h: pcfa_corr_occupancy_mark)
            occ = G[st.total:]
            for radius in sorted({r for _, r in lookups}):
                ptrs = [c.data_ptr() for c, r in lookups if r == radius]
                arr = (C.c_void_p * len(ptrs))(*ptrs)
                _lib.check(lib.pcfa_corr_occupancy_mark(arr, len(ptrs), _lib.ptr(occ), st.B, st.H, st.W, st.levels, radius,
                                                        _lib.stream()), "pcfa_corr_occupancy_mark")
        g1 = torch.empty_like(fmap1)
        g2 = torch.empty_like(fmap2)
        wsb = lib.pcfa_corr_pyramid_workspace_bytes(st.B, st.C, st.H, st.W, st.levels)
        ws = torch.empty(wsb, device=fmap1.device, dtype=torch.uint8)
        if occ is not None:
            _lib.check(lib.pcfa_corr_pyramid_backward_occ(_lib.ptr(G), _lib.ptr(occ), _lib.ptr(fmap1), _lib.ptr(fmap2),
                                                          _lib.ptr(g1), _lib.ptr(g2), _lib.ptr(ws), wsb,
                                                          st.B, st.C, st.H, st.W, st.levels, st.impl,
                                                          _lib.stream()), "pcfa_corr_pyramid_backward_occ")
        else:
            _lib.check(lib.pcfa_corr_pyramid_backward(_lib.ptr(G), _lib.ptr(fmap1), _lib.ptr(fmap2),
                                                      _lib.ptr(g1), _lib.ptr(g2), _lib.ptr(ws), wsb,
                                                      st.B, st.C, st.H, st.W, st.levels, st.impl,
                                                      _lib.stream()), "pcfa_corr_pyramid_backward")
        return g1, g2, None


class _Lookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pyr, coords, state, radius, channels_last=False):
        lib = _lib.load()
        _lib.require_cuda(pyr, coords, name="CorrBlock.__call__")
        B, H, W, L = state.B, state.H, state.W, state.levels
        D = 2 * radius + 1
        fmt = torch.channels_last if channels_last else torch.contiguous_format
        out = torch.empty((B, L * D * D, H, W), device=pyr.device, dtype=torch.float32, memory_format=fmt)
        fn = lib.pcfa_corr_lookup_forward_cl if channels_last else lib.pcfa_corr_lookup_forward
        _lib.check(fn(_lib.ptr(pyr), _lib.ptr(coords), _lib.ptr(out), B, H, W, L, radius, _lib.stream()),
                   "pcfa_corr_lookup_forward")
        ctx.save_for_backward(coords)
        ctx.state, ctx.radius, ctx.cl = state, radius, bool(channels_last)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        (coords,) = ctx.saved_tensors
        st = ctx.state
        if st.grad is None:
            st.new_grad(gout.device)
        gout = gout.contiguous(memory_format=torch.channels_last) if ctx.cl else gout.contiguous()
        fn = lib.pcfa_corr_lookup_backward_cl if ctx.cl else lib.pcfa_corr_lookup_backward
        _lib.check(fn(_lib.ptr(gout), _lib.ptr(coords), _lib.ptr(st.grad), st.B, st.H, st.W, st.levels, ctx.radius,
                      _lib.stream()), "pcfa_corr_lookup_backward")
        st.lookups.append((coords, ctx.radius))
        # the pyramid's gradient travels through st.grad (consumed by _BuildPyramid.backward, which
        # autograd runs after every lookup node); coords are detached in RAFT/GMA (raft.py:123)
        return None, None, None, None, None


class CorrBlock:
    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        self.num_levels = num_levels
        self.radius = radius
        fmap1 = fmap1.float().contiguous()
        fmap2 = fmap2.float().contiguous()
        B, Cc, H, W = fmap1.shape
        offs, hs, ws = pyramid_layout(B, H, W, num_levels)
        self._state = _PyramidState(B, Cc, H, W, num_levels, offs[-1], _impl_from_env())
        self._flat = _BuildPyramid.apply(fmap1, fmap2, self._state)
        N = B * H * W
        self.corr_pyramid = [self._flat[offs[l]:offs[l + 1]].view(N, 1, hs[l], ws[l])
                             for l in range(num_levels)]

    def __call__(self, coords, channels_last=False):
        """[B, levels*(2r+1)^2, H, W]; channels_last=True returns the same values in torch.channels_last memory
        (what an NHWC update block consumes without a layout conversion)."""
        coords = coords.detach().float().contiguous()
        return _Lookup.apply(self._flat, coords, self._state, self.radius, channels_last)

    @staticmethod
    def corr(fmap1, fmap2):
        fmap1 = fmap1.float().contiguous()
        fmap2 = fmap2.float().contiguous()
        B, Cc, H, W = fmap1.shape
        offs, _, _ = pyramid_layout(B, H, W, 1)
        st = _PyramidState(B, Cc, H, W, 1, offs[-1], _impl_from_env())
        return _BuildPyramid.apply(fmap1, fmap2, st).view(B, H, W, 1, H, W)
