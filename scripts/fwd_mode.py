"""Forward all-pairs pyramid: arithmetic mode (PCFA_FWD_MODE 0 = bf16x3, 1 = fp16x2) vs the exact-fp32 SIMT build at
1x256x55x128 — device time (events, L2 flushed) and error against BASELINE's element-wise tolerance."""
import json, os, statistics, sys
import numpy as np
import torch
sys.path.insert(0, '.')
from pcfa_b200 import _lib
from pcfa_b200.corr_block import pyramid_layout
lib = _lib.load(); P = _lib.ptr; s = _lib.stream()
B, C, H, W, L = 1, 256, 55, 128, 4
kind = os.environ.get("FEAT", "randn")
g = torch.Generator().manual_seed(0)
f1 = torch.randn(B, C, H, W, generator=g).cuda(); f2 = torch.randn(B, C, H, W, generator=g).cuda()
if kind == "corr":            # correlated features (shifted copy + noise): strong peaks like a real cost volume
    f2 = (torch.roll(f1, (2, 3), (2, 3)) + 0.3 * f2).contiguous()
offs, hs, ws = pyramid_layout(B, H, W, L)
wsb = lib.pcfa_corr_pyramid_workspace_bytes(B, C, H, W, L); wsp = torch.empty(wsb, device="cuda", dtype=torch.uint8)
ref = torch.empty(offs[-1], device="cuda"); out = torch.empty(offs[-1], device="cuda")
assert lib.pcfa_corr_pyramid_forward(P(f1), P(f2), P(ref), P(wsp), wsb, B, C, H, W, L, 1, s) == 0     # fp32 SIMT
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(18):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st = lib.pcfa_corr_pyramid_forward(P(f1), P(f2), P(out), P(wsp), wsb, B, C, H, W, L, 0, s); e1.record(); torch.cuda.synchronize()
    assert st == 0
    if i >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
d = (out - ref).double(); r = ref.double()
rms = float(r.pow(2).mean().sqrt())
viol = (d.abs() - 1e-3 * r.abs()) / rms
res = dict(mode=os.environ.get("PCFA_FWD_MODE", "0"), feat=kind, us=round(statistics.median(ts), 1), rel_l2=float(d.norm() / r.norm()),
           max_abs_over_rms=float(d.abs().max() / rms), max_violation_over_rms=float(viol.max()),
           passes_rtol1e3_atolrms1e3=bool(viol.max() <= 1e-3), frac_above_5e4rms=float((d.abs() > 5e-4 * rms).double().mean()))
for l in range(L):
    dl, rl = d[offs[l]:offs[l + 1]], r[offs[l]:offs[l + 1]]
    res[f"rel_l2_level{l}"] = float(dl.norm() / rl.norm())
print(json.dumps(res))
