"""Per-kernel timing at the BASELINE shapes (RAFT 440x1024 -> 55x128 features, C=256, 4 levels, r=4):
CUDA events on the launching stream, L2 flushed between repetitions, median of N.  Not the headline
bench (that is bench.py); this is the inner loop for kernel tuning."""
import json, statistics, sys
import torch
sys.path.insert(0, '.')
from pcfa_b200 import _lib
from pcfa_b200.corr_block import pyramid_layout
from pcfa_b200.profiling import algorithmic_bytes

lib = _lib.load()
B, C, H, W, L, R = int(sys.argv[1]) if len(sys.argv) > 1 else 1, 256, 55, 128, 4, 4
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = 15
noflush = len(sys.argv) > 3 and sys.argv[3] == 'noflush'   # warm-L2 variant: what the kernel does when its inputs are L2-resident
g = torch.Generator().manual_seed(0)
f1 = torch.randn(B, C, H, W, generator=g).cuda(); f2 = torch.randn(B, C, H, W, generator=g).cuda()
ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
coords = (torch.stack([xs, ys]).float()[None] + 3 * torch.randn(B, 2, H, W, generator=g)).cuda().contiguous()
offs, hs, ws = pyramid_layout(B, H, W, L)
pyr = torch.empty(offs[-1], device="cuda"); gpyr = torch.zeros(offs[-1], device="cuda")
wsb = lib.pcfa_corr_pyramid_workspace_bytes(B, C, H, W, L); wsp = torch.empty(wsb, device="cuda", dtype=torch.uint8)
out = torch.empty(B, L * 81, H, W, device="cuda"); gout = torch.randn(B, L * 81, H, W, device="cuda")
g1 = torch.empty_like(f1); g2 = torch.empty_like(f2)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
s = _lib.stream()
P = _lib.ptr
calls = {
    "pcfa_corr_pyramid_forward": lambda: lib.pcfa_corr_pyramid_forward(P(f1), P(f2), P(pyr), P(wsp), wsb, B, C, H, W, L, impl, s),
    "pcfa_corr_lookup_forward": lambda: lib.pcfa_corr_lookup_forward(P(pyr), P(coords), P(out), B, H, W, L, R, s),
    "pcfa_corr_lookup_backward": lambda: lib.pcfa_corr_lookup_backward(P(gout), P(coords), P(gpyr), B, H, W, L, R, s),
    "pcfa_corr_pyramid_backward": lambda: lib.pcfa_corr_pyramid_backward(P(gpyr), P(f1), P(f2), P(g1), P(g2), P(wsp), wsb, B, C, H, W, L, impl, s),
}
peak = 6550.4
rows = []
for name, fn in calls.items():
    ts = []
    for i in range(reps + 3):
        if not noflush: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); st = fn(); e1.record(); torch.cuda.synchronize()
        assert st == 0, (name, st)
        if i >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
    ab = algorithmic_bytes(name, B=B, C=C, H=H, W=W)
    us = statistics.median(ts)
    rows.append(dict(name=name, us=round(us, 1), min_us=round(min(ts), 1), gbs=round(ab / us / 1e3, 1), frac=round(ab / us / 1e3 / peak, 3)))
    print(rows[-1])
json.dump(rows, open(("gpurun_out/bench_kernels.json" if B == 1 else f"gpurun_out/bench_kernels_B{B}.json") if not noflush else f"gpurun_out/bench_kernels_B{B}_warm.json", "w"))
