import statistics, sys, torch
sys.path.insert(0, '.')
from pcfa_b200 import _lib
from pcfa_b200.corr_block import pyramid_layout
lib = _lib.load(); P = _lib.ptr; s = _lib.stream()
B, C, H, W, L = 1, 256, 55, 128, 4
f1 = torch.randn(B, C, H, W).cuda(); f2 = torch.randn(B, C, H, W).cuda()
offs, _, _ = pyramid_layout(B, H, W, L)
pyr = torch.empty(offs[-1], device="cuda")
wsb = lib.pcfa_corr_pyramid_workspace_bytes(B, C, H, W, L); wsp = torch.empty(wsb, device="cuda", dtype=torch.uint8)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(13):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st = lib.pcfa_corr_pyramid_forward(P(f1), P(f2), P(pyr), P(wsp), wsb, B, C, H, W, L, 0, s); e1.record(); torch.cuda.synchronize()
    assert st == 0
    if i >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
print(sys.argv[1:], "fwd us median", round(statistics.median(ts), 1), "min", round(min(ts), 1))
