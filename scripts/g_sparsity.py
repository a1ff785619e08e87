"""How sparse is the gradient pyramid G that the cost-volume backward contracts?  Runs the bench closure's forward,
records every lookup's coords, scatters ones through pcfa_corr_lookup_backward_cl and reports the fraction of non-zero
elements and of non-zero GEMM tiles (pass I: 512 queries x 32 cells; pass II: 512 cells x 32 queries)."""
import json, sys
import torch
sys.path.insert(0, '.')
from pcfa_b200 import _lib, corr_block
from pcfa_b200.adapter import build_network, preprocess_img
from pcfa_b200.networks.weights import synthetic_pair
lib = _lib.load()
gain = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
net = build_network("RAFT", device="cuda", seed=0, gain=gain)
i1, i2 = synthetic_pair(0, 436, 1024)
_, (a, b) = preprocess_img("RAFT", i1.cuda(), i2.cuda())
seen = []
orig = corr_block.CorrBlock.__call__
def spy(self, coords, channels_last=False):
    seen.append(coords.detach().clone()); return orig(self, coords, channels_last)
corr_block.CorrBlock.__call__ = spy
with torch.no_grad():
    net(a.contiguous(), b.contiguous(), iters=12, test_mode=True)
B, H, W, L, R = 1, 55, 128, 4, 4
offs, hs, ws = corr_block.pyramid_layout(B, H, W, L)
G = torch.zeros(offs[-1], device="cuda")
go = torch.ones(B, H, W, L * 81, device="cuda")
for c in seen:
    assert lib.pcfa_corr_lookup_backward_cl(_lib.ptr(go), _lib.ptr(c.contiguous()), _lib.ptr(G), B, H, W, L, R, _lib.stream()) == 0
res = {"gain": gain, "lookups": len(seen), "nonzero_fraction": float((G != 0).float().mean())}
N = H * W
tot1 = nz1 = tot2 = nz2 = 0
for l in range(L):
    nl = hs[l] * ws[l]
    g = (G[offs[l]:offs[l + 1]].view(N, nl) != 0)
    qp, cp = (-N) % 512, (-nl) % 32
    t = torch.nn.functional.pad(g, (0, cp, 0, qp)).view((N + qp) // 512, 512, (nl + cp) // 32, 32).any(3).any(1)
    res[f"passI_level{l}"] = float(t.float().mean()); tot1 += t.numel() * (1 if True else 0); nz1 += int(t.sum())
    qp, cp = (-N) % 32, (-nl) % 512
    t2 = torch.nn.functional.pad(g, (0, cp, 0, qp)).view((N + qp) // 32, 32, (nl + cp) // 512, 512).any(3).any(1)
    res[f"passII_level{l}"] = float(t2.float().mean()); tot2 += t2.numel(); nz2 += int(t2.sum())
    # finer tiles: 128 queries x 32 cells
    qp, cp = (-N) % 128, (-nl) % 32
    t3 = torch.nn.functional.pad(g, (0, cp, 0, qp)).view((N + qp) // 128, 128, (nl + cp) // 32, 32).any(3).any(1)
    res[f"tile128x32_level{l}"] = float(t3.float().mean())
res["passI_tiles_nonzero"] = nz1 / tot1; res["passII_tiles_nonzero"] = nz2 / tot2
fl = (seen[-1] - seen[0])
res["flow_cells_abs_mean"] = float(fl.abs().mean()); res["flow_cells_abs_max"] = float(fl.abs().max())
print(json.dumps(res))
