"""Timing of the local-window correlation, warp, channelnorm and objective kernels at the BASELINE shapes
(PWCNet 384x1280 pyramid levels, FlowNet2 48x160 C=256 correlation, 384x1280 warps), CUDA events, L2 flushed."""
import ctypes as C, json, statistics, sys
import torch
sys.path.insert(0, '.')
from pcfa_b200 import _lib
lib = _lib.load(); P = _lib.ptr; s = _lib.stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
PEAK = 6550.4
rows = []

def timeit(name, fn, bytes_, flops=0, reps=12):
    ts = []
    for i in range(reps + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); st = fn(); e1.record(); torch.cuda.synchronize()
        assert st == 0, (name, st)
        if i >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
    us = statistics.median(ts)
    r = dict(name=name, us=round(us, 1), algorithmic_MB=round(bytes_ / 1e6, 2), gbs=round(bytes_ / us / 1e3, 1),
             frac_hbm=round(bytes_ / us / 1e3 / PEAK, 3))
    if flops: r["gflops"] = round(flops / us / 1e3, 1)
    rows.append(r); print(r)

g = torch.Generator().manual_seed(0)
# ---- PWCNet correlations (C,H,W) per level, B=1, patch 9
for (Cc, H, W) in [(196, 6, 20), (128, 12, 40), (96, 24, 80), (64, 48, 160), (32, 96, 320)]:
    a = torch.randn(1, Cc, H, W, generator=g).cuda(); b = torch.randn(1, Cc, H, W, generator=g).cuda()
    out = torch.empty(1, 9, 9, H, W, device="cuda"); go = torch.randn_like(out); g1 = torch.empty_like(a); g2 = torch.empty_like(b)
    p = _lib.ScsParams(1, 1, 9, 9, 0, 0, 1, 1, 1, 1, 1, 1)
    inb, outb = 2 * a.numel() * 4, out.numel() * 4
    timeit(f"scs_fwd C{Cc} {H}x{W}", lambda: lib.pcfa_scs_forward(P(a), P(b), P(out), 1, Cc, H, W, C.byref(p), 1.0 / Cc, s), inb + outb, 2 * 81 * Cc * H * W)
    timeit(f"scs_bwd C{Cc} {H}x{W}", lambda: lib.pcfa_scs_backward(P(a), P(b), P(go), P(g1), P(g2), 1, Cc, H, W, C.byref(p), 1.0 / Cc, s), 2 * inb + outb, 4 * 81 * Cc * H * W)
# ---- FlowNet2 correlation C=256 48x160, 21x21 stride2 2
a = torch.randn(1, 256, 48, 160, generator=g).cuda(); b = torch.randn(1, 256, 48, 160, generator=g).cuda()
out = torch.empty(1, 441, 48, 160, device="cuda"); go = torch.randn_like(out); g1 = torch.empty_like(a); g2 = torch.empty_like(b)
timeit("fn2corr_fwd", lambda: lib.pcfa_fn2corr_forward(P(a), P(b), P(out), 1, 256, 48, 160, 20, 1, 20, 1, 2, s), 2 * a.numel() * 4 + out.numel() * 4, 2 * 441 * 256 * 48 * 160)
timeit("fn2corr_bwd", lambda: lib.pcfa_fn2corr_backward(P(a), P(b), P(go), P(g1), P(g2), 1, 256, 48, 160, 20, 1, 20, 1, 2, s), 4 * a.numel() * 4 + out.numel() * 4, 4 * 441 * 256 * 48 * 160)
# ---- Resample2d / ChannelNorm 384x1280
img = torch.randn(1, 3, 384, 1280, generator=g).cuda(); flow = (3 * torch.randn(1, 2, 384, 1280, generator=g)).cuda()
o = torch.empty_like(img); gi = torch.zeros_like(img); gf = torch.empty_like(flow); go = torch.randn_like(img)
timeit("resample2d_fwd", lambda: lib.pcfa_resample2d_forward(P(img), P(flow), P(o), 1, 3, 384, 1280, 384, 1280, 1, 1, s), (3 + 2 + 3) * 384 * 1280 * 4)
timeit("resample2d_bwd", lambda: lib.pcfa_resample2d_backward(P(img), P(flow), P(go), P(gi), P(gf), 1, 3, 384, 1280, 384, 1280, 1, 1, s), (3 + 2 + 3 + 2 * 3 + 2) * 384 * 1280 * 4)
n = torch.empty(1, 1, 384, 1280, device="cuda"); gn = torch.randn_like(n); gx = torch.empty_like(img)
timeit("channelnorm_fwd", lambda: lib.pcfa_channelnorm_forward(P(img), P(n), 1, 3, 384, 1280, 2, s), 4 * 384 * 1280 * 4)
timeit("channelnorm_bwd", lambda: lib.pcfa_channelnorm_backward(P(img), P(n), P(gn), P(gx), 1, 3, 384, 1280, 2, s), 8 * 384 * 1280 * 4)
# ---- PWC warp level 2 (C=32, 96x320) and level 5 (128, 12x40)
for (Cc, H, W) in [(32, 96, 320), (128, 12, 40)]:
    x = torch.randn(1, Cc, H, W, generator=g).cuda(); f = (2 * torch.randn(1, 2, H, W, generator=g)).cuda()
    o = torch.empty_like(x); go = torch.randn_like(x); gx = torch.zeros_like(x); gf = torch.empty_like(f)
    timeit(f"pwc_warp_fwd C{Cc} {H}x{W}", lambda: lib.pcfa_pwc_warp_forward(P(x), P(f), P(o), 1, Cc, H, W, s), (2 * Cc + 2) * H * W * 4)
    timeit(f"pwc_warp_bwd C{Cc} {H}x{W}", lambda: lib.pcfa_pwc_warp_backward(P(x), P(f), P(go), P(gx), P(gf), 1, Cc, H, W, s), (4 * Cc + 4) * H * W * 4)
json.dump(rows, open("gpurun_out/bench_ops.json", "w"), indent=1)
