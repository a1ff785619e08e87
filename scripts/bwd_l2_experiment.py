"""Is the TF32 backward limited by DRAM (latency / row locality) or by the TMA/MMA pipeline?  Times the backward
entry point on a gradient pyramid that fits L2, cold (L2 flushed) vs warm (same launch repeated)."""
import statistics, sys, torch
sys.path.insert(0, '.')
from pcfa_b200 import _lib
from pcfa_b200.corr_block import pyramid_layout
lib = _lib.load(); P = _lib.ptr; s = _lib.stream()
for (B, H, W) in ((2, 46, 64), (1, 55, 128)):
    C, L = 256, 4
    f1 = torch.randn(B, C, H, W).cuda(); f2 = torch.randn(B, C, H, W).cuda()
    offs, _, _ = pyramid_layout(B, H, W, L)
    gp = torch.randn(offs[-1], device="cuda")
    wsb = lib.pcfa_corr_pyramid_workspace_bytes(B, C, H, W, L); wsp = torch.empty(wsb, device="cuda", dtype=torch.uint8)
    g1 = torch.empty_like(f1); g2 = torch.empty_like(f2)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for mode in ("cold", "warm"):
        ts = []
        for i in range(12):
            if mode == "cold": flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); st = lib.pcfa_corr_pyramid_backward(P(gp), P(f1), P(f2), P(g1), P(g2), P(wsp), wsb, B, C, H, W, L, 0, s); e1.record()
            torch.cuda.synchronize(); assert st == 0
            if i >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
        print((B, H, W), "G MB", round(offs[-1] * 4 / 1e6, 1), mode, "us", round(statistics.median(ts), 1), flush=True)
