"""Runs the all-pairs pyramid build (and optionally lookups / backward) at the BASELINE shape — the
short command ncu captures are taken from."""
import sys, torch
sys.path.insert(0, '.')
from pcfa_b200 import _lib
from pcfa_b200.corr_block import CorrBlock
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
what = sys.argv[2] if len(sys.argv) > 2 else "fwd"
g = torch.Generator().manual_seed(0)
f1 = torch.randn(B, 256, 55, 128, generator=g).cuda().requires_grad_(True)
f2 = torch.randn(B, 256, 55, 128, generator=g).cuda().requires_grad_(True)
ys, xs = torch.meshgrid(torch.arange(55), torch.arange(128), indexing="ij")
coords = (torch.stack([xs, ys]).float()[None] + 3 * torch.randn(B, 2, 55, 128, generator=g)).cuda()
for it in range(3):
    blk = CorrBlock(f1, f2)
    if what != "fwd":
        outs = [blk(coords + 0.1 * i) for i in range(2)]
        sum(o.sum() for o in outs).backward()
torch.cuda.synchronize()
print("done", _lib.launch_count())
