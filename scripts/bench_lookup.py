"""Lookup kernels (channels-last entry points) at RAFT's BASELINE shape, 1x55x128, 4 levels, r=4.
(a) isolated: L2 flushed before every launch; (b) in sequence: 12 launches with slowly drifting coords and ~150 MB of
unrelated streaming traffic (a stand-in for one GRU iteration's activations) between them, no flush.
Variant selection by environment: PCFA_LOOKUP_IMPL (1 = first generation, 2 = warp-per-task), PCFA_LOOKUP_L2HINT."""
import json, os, statistics, sys
import torch
sys.path.insert(0, '.')
from pcfa_b200 import _lib
lib = _lib.load(); P = _lib.ptr; s = _lib.stream()
B = int(os.environ.get("B", "1")); H, W, L, R = 55, 128, 4, 4
N = H * W
tot = sum(B * N * (H >> l) * (W >> l) for l in range(L))
g = torch.Generator(device="cuda").manual_seed(0)
pyr = torch.randn(tot, device="cuda", generator=g)
gp = torch.zeros(tot, device="cuda")
ys, xs = torch.meshgrid(torch.arange(H, device="cuda", dtype=torch.float32), torch.arange(W, device="cuda", dtype=torch.float32), indexing="ij")
base = torch.stack([xs, ys])[None].repeat(B, 1, 1, 1)
flow = 6 * torch.randn(B, 2, H, W, device="cuda", generator=g)
flow = torch.nn.functional.avg_pool2d(flow, 9, 1, 4)          # smooth flow field, a few cells
out = torch.empty(B, N, L * 81, device="cuda"); go = torch.randn_like(out)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream_a = torch.empty(75 << 18, device="cuda"); stream_b = torch.empty_like(stream_a)    # 75 MB each: copy = 150 MB traffic

def ev(): return torch.cuda.Event(enable_timing=True)
def bracket(fn):
    e0, e1 = ev(), ev(); e0.record(); st = fn(); e1.record(); torch.cuda.synchronize(); assert st == 0; return e0.elapsed_time(e1) * 1e3
empty = statistics.median(bracket(lambda: 0) for _ in range(30))
res = {"impl": os.environ.get("PCFA_LOOKUP_IMPL", "2"), "hint": os.environ.get("PCFA_LOOKUP_L2HINT", "1"), "B": B, "empty_us": round(empty, 2)}
for name, fn_ in (("fwd", lambda c: lib.pcfa_corr_lookup_forward_cl(P(pyr), P(c), P(out), B, H, W, L, R, s)),
                  ("bwd", lambda c: lib.pcfa_corr_lookup_backward_cl(P(go), P(c), P(gp), B, H, W, L, R, s))):
    c0 = (base + flow).contiguous()
    ts = []
    for i in range(15):
        flush.zero_(); t = bracket(lambda: fn_(c0))
        if i >= 3: ts.append(t - empty)
    res[name + "_isolated_us"] = round(statistics.median(ts), 2)
    seq = []
    for rep in range(4):
        flush.zero_()
        for it in range(12):
            c = (base + flow * (1 + 0.02 * it) + 0.05 * it).contiguous()
            stream_b.copy_(stream_a)
            torch.cuda._sleep(20000)
            t = bracket(lambda: fn_(c))
            if rep >= 1 and it >= 1: seq.append(t - empty)
    res[name + "_sequence_us"] = round(statistics.median(seq), 2)
print(json.dumps(res))
