"""Per-CTA timeline of the cta_group::2 backward kernel (globaltimer stamps, see pcfa_debug_set_bwd_trace)."""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, '.')
from pcfa_b200 import _lib
from pcfa_b200.corr_block import pyramid_layout
lib = _lib.load(); P = _lib.ptr; s = _lib.stream()
raw = ctypes.CDLL(str(_lib.lib_path()))
raw.pcfa_debug_set_bwd_trace.argtypes = [ctypes.c_void_p]
B, C, H, W, L = int(sys.argv[1]) if len(sys.argv) > 1 else 1, 256, 55, 128, 4
f1 = torch.randn(B, C, H, W).cuda(); f2 = torch.randn(B, C, H, W).cuda()
offs, _, _ = pyramid_layout(B, H, W, L)
gp = torch.randn(offs[-1], device="cuda")
wsb = lib.pcfa_corr_pyramid_workspace_bytes(B, C, H, W, L); wsp = torch.empty(wsb, device="cuda", dtype=torch.uint8)
g1 = torch.empty_like(f1); g2 = torch.empty_like(f2)
trace = torch.zeros(2 * 1024 * 8, dtype=torch.int64, device="cuda")
for it in range(3):
    trace.zero_()
    raw.pcfa_debug_set_bwd_trace(ctypes.c_void_p(trace.data_ptr()))
    assert lib.pcfa_corr_pyramid_backward(P(gp), P(f1), P(f2), P(g1), P(g2), P(wsp), wsb, B, C, H, W, L, 0, s) == 0
    torch.cuda.synchronize()
raw.pcfa_debug_set_bwd_trace(None)
t = trace.cpu().numpy().reshape(2, 1024, 8)[:, :148].astype(np.float64)
names = ["entry", "setup done", "first data at MMA", "last MMA issued", "epi: accfull (not last)", "epi: last accfull", "epi: reductions done", "after cluster sync"]
for p in range(2):
    tp = t[p]; t0 = tp[:, 0][tp[:, 0] > 0].min()
    print(f"pass {p + 1}: stamps relative to the earliest CTA entry, us (min / median / max over CTAs that wrote the slot)")
    for k, n in enumerate(names):
        v = tp[:, k]; v = v[v > 0]
        if len(v): print(f"  {k} {n:28s} n={len(v):3d}  {(v.min() - t0) / 1e3:7.2f} {(np.median(v) - t0) / 1e3:7.2f} {(v.max() - t0) / 1e3:7.2f}")
