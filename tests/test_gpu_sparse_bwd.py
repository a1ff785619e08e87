"""Sparse backward of the correlation pyramid (pcfa_corr_occupancy_mark + pcfa_corr_pyramid_backward_occ): the occupancy
bitmap marked from the lookups' coordinates is a superset of the non-zero 32x32 blocks of the gradient pyramid for every
lookup kernel and radius, and the block-skipping build backward returns what the dense one returns on the same input."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _occ_layout(H, W, L):
    qg = (H * W + 31) // 32
    off, h, w = [], H, W
    c = 0
    for _ in range(L):
        off.append(c)
        c += (h * w + 31) // 32
        h, w = h // 2, w // 2
    off.append(c)
    return qg, off, (c + 31) // 32


def _coords(B, H, W, spread, seed):
    g = np.random.default_rng(seed)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    c = (np.stack([xs, ys])[None] + spread * g.standard_normal((B, 2, H, W))).astype(np.float32)
    c[0, :, 0, :4] = np.array([[-30.0, W - 1.0, 0.0, 500.0], [2.0, H - 1.0, 0.0, -7.5]], np.float32)   # out-of-range windows
    return torch.from_numpy(c).cuda()


def _run_lookups(B, H, W, L, R, coords_list, gouts, occ, cl=True):
    import ctypes
    from oracle import ops as O
    from pcfa_b200 import _lib
    lib = _lib.load()
    off, _, _ = O.pyramid_layout(B, H, W, L)
    nocc = lib.pcfa_corr_occupancy_bytes(B, H, W, L) // 4
    buf = torch.zeros(off[-1] + nocc, device="cuda")
    fn = lib.pcfa_corr_lookup_backward_cl if cl else lib.pcfa_corr_lookup_backward
    for c, go in zip(coords_list, gouts):
        _lib.check(fn(_lib.ptr(go), _lib.ptr(c), _lib.ptr(buf), B, H, W, L, R, _lib.stream()), "lookup bwd")
    if occ:
        arr = (ctypes.c_void_p * len(coords_list))(*[c.data_ptr() for c in coords_list])
        _lib.check(lib.pcfa_corr_occupancy_mark(arr, len(coords_list), _lib.ptr(buf[off[-1]:]), B, H, W, L, R, _lib.stream()), "mark")
    return buf, off, nocc


@pytest.mark.parametrize("B,H,W,R,cl", [(1, 55, 128, 4, True), (2, 46, 62, 4, True), (1, 47, 156, 3, True), (1, 40, 90, 2, False)])
def test_occupancy_bitmap_covers_the_gradient_of_every_lookup_kernel(B, H, W, R, cl):
    L, D = 4, 2 * R + 1
    g = torch.Generator(device="cuda").manual_seed(3)
    coords = [_coords(B, H, W, 5.0, s) for s in (1, 2)]
    fmt = torch.channels_last if cl else torch.contiguous_format
    gouts = [torch.randn((B, L * D * D, H, W), device="cuda", generator=g).contiguous(memory_format=fmt) for _ in coords]
    marked, off, nocc = _run_lookups(B, H, W, L, R, coords, gouts, occ=True, cl=cl)
    qg, coff, words = _occ_layout(H, W, L)
    assert nocc == B * qg * words
    bits = marked[off[-1]:].view(torch.int32).cpu().numpy().view(np.uint32).reshape(B, qg, words)
    N, h, w = H * W, H, W
    total_blocks = marked_blocks = 0
    for l in range(L):
        nl = h * w
        G = marked[off[l]:off[l + 1]].view(B, N, nl)
        nz = (G != 0)
        padq, padc = qg * 32 - N, (coff[l + 1] - coff[l]) * 32 - nl
        nz = torch.nn.functional.pad(nz, (0, padc, 0, padq))
        blk = nz.view(B, qg, 32, coff[l + 1] - coff[l], 32).any(dim=4).any(dim=2).cpu().numpy()     # [B, qg, chunks_l]
        cols = np.arange(coff[l], coff[l + 1])
        have = ((bits[:, :, cols >> 5] >> (cols & 31).astype(np.uint32)) & 1).astype(bool)
        assert not (blk & ~have).any(), "level %d: a non-zero block is not marked" % l
        total_blocks += blk.size
        marked_blocks += int(have.sum())
        h, w = h // 2, w // 2
    assert 0 < marked_blocks < total_blocks                                     # it is a real subset at these spreads


@pytest.mark.parametrize("B,C,H,W,spread", [(1, 256, 55, 128, 3.0), (1, 256, 55, 128, 40.0), (2, 128, 48, 64, 4.0), (1, 64, 32, 96, 2.0), (2, 128, 46, 62, 4.0)])
def test_sparse_pyramid_backward_equals_dense(B, C, H, W, spread):
    from pcfa_b200 import _lib
    lib = _lib.load()
    L, R, D = 4, 4, 9
    g = torch.Generator(device="cuda").manual_seed(11)
    coords = [_coords(B, H, W, spread, s) for s in (5, 6, 7)]
    gouts = [torch.randn((B, L * D * D, H, W), device="cuda", generator=g).contiguous(memory_format=torch.channels_last) for _ in coords]
    buf, off, nocc = _run_lookups(B, H, W, L, R, coords, gouts, occ=True)
    f1 = torch.randn((B, C, H, W), device="cuda", generator=g)
    f2 = torch.randn((B, C, H, W), device="cuda", generator=g)
    wsb = lib.pcfa_corr_pyramid_workspace_bytes(B, C, H, W, L)
    ws = torch.empty(wsb, device="cuda", dtype=torch.uint8)
    res = []
    for sparse in (False, True):
        g1, g2 = torch.full_like(f1, float("nan")), torch.full_like(f2, float("nan"))
        if sparse:
            _lib.check(lib.pcfa_corr_pyramid_backward_occ(_lib.ptr(buf), _lib.ptr(buf[off[-1]:]), _lib.ptr(f1), _lib.ptr(f2), _lib.ptr(g1),
                                                          _lib.ptr(g2), _lib.ptr(ws), wsb, B, C, H, W, L, 0, _lib.stream()), "bwd occ")
        else:
            _lib.check(lib.pcfa_corr_pyramid_backward(_lib.ptr(buf), _lib.ptr(f1), _lib.ptr(f2), _lib.ptr(g1), _lib.ptr(g2),
                                                      _lib.ptr(ws), wsb, B, C, H, W, L, 0, _lib.stream()), "bwd")
        torch.cuda.synchronize()
        res.append((g1.clone(), g2.clone()))
    for a, b, what in ((res[0][0], res[1][0], "grad_fmap1"), (res[0][1], res[1][1], "grad_fmap2")):
        assert torch.isfinite(b).all(), what
        # the same products in a different split-K order: fp32 summation noise only
        err = float((a - b).abs().max()) / float(a.abs().max())
        assert err < 2e-5, (what, err)


def test_sparse_pyramid_backward_with_empty_bitmap_returns_zeros():
    from oracle import ops as O
    from pcfa_b200 import _lib
    lib = _lib.load()
    B, C, H, W, L = 1, 256, 55, 128, 4
    off, _, _ = O.pyramid_layout(B, H, W, L)
    nocc = lib.pcfa_corr_occupancy_bytes(B, H, W, L) // 4
    buf = torch.zeros(off[-1] + nocc, device="cuda")
    f1, f2 = torch.randn((B, C, H, W), device="cuda"), torch.randn((B, C, H, W), device="cuda")
    g1, g2 = torch.full_like(f1, float("nan")), torch.full_like(f2, float("nan"))
    wsb = lib.pcfa_corr_pyramid_workspace_bytes(B, C, H, W, L)
    ws = torch.empty(wsb, device="cuda", dtype=torch.uint8)
    _lib.check(lib.pcfa_corr_pyramid_backward_occ(_lib.ptr(buf), _lib.ptr(buf[off[-1]:]), _lib.ptr(f1), _lib.ptr(f2), _lib.ptr(g1),
                                                  _lib.ptr(g2), _lib.ptr(ws), wsb, B, C, H, W, L, 0, _lib.stream()), "bwd occ")
    assert float(g1.abs().max()) == 0.0 and float(g2.abs().max()) == 0.0


def test_corr_block_autograd_sparse_matches_dense_and_oracle():
    """CorrBlock end to end (build, 3 channels-last lookups, backward) with the sparse path (default) and with
    PCFA_BWD_SPARSE=0 in a second process: same gradients up to summation order."""
    code = r'''
import sys, torch
from pcfa_b200.corr_block import CorrBlock
torch.manual_seed(0)
B, C, H, W = 1, 256, 55, 128
f1 = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
f2 = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
ys, xs = torch.meshgrid(torch.arange(H, device="cuda"), torch.arange(W, device="cuda"), indexing="ij")
base = torch.stack([xs, ys]).float()[None]
cb = CorrBlock(f1, f2, num_levels=4, radius=4)
loss = 0
for i in range(3):
    out = cb(base + 3.0 * torch.randn(B, 2, H, W, device="cuda"), channels_last=True)
    loss = loss + (out * torch.randn_like(out)).sum()
loss.backward()
torch.save({"g1": f1.grad.cpu(), "g2": f2.grad.cpu()}, sys.argv[1])
'''
    import tempfile
    outs = []
    with tempfile.TemporaryDirectory() as td:
        for tag, env in (("sparse", {"PCFA_BWD_SPARSE": "1"}), ("dense", {"PCFA_BWD_SPARSE": "0"})):
            path = os.path.join(td, tag + ".pt")
            r = subprocess.run([sys.executable, "-c", code, path], env=dict(os.environ, **env), capture_output=True, text=True,
                               cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(torch.load(path))
    for k in ("g1", "g2"):
        a, b = outs[0][k], outs[1][k]
        err = float((a - b).abs().max()) / float(b.abs().max())
        assert err < 2e-5, (k, err)
