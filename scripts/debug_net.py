import sys, numpy as np, torch
sys.path.insert(0, '.')
from oracle import torch_ref as TR
from pcfa_b200.adapter import build_network, preprocess_img
from pcfa_b200.networks.weights import synthetic_pair
from pcfa_b200 import objective as J

def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30)), float((a - b).norm() / (b.norm() + 1e-30))

for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    print("=== cudnn tf32", tf32)
    gout = torch.randn(1, 2, 128, 160, generator=torch.Generator().manual_seed(15)).cuda() / (2 * 128 * 160)
    for iters in (1, 2, 4, 12):
        res = {}
        for tag, ops in (("cuda", None), ("torch", TR), ("cuda2", None), ("torch2", TR)):
            net = build_network("RAFT", device="cuda", seed=0, ops=ops)
            i1, i2 = synthetic_pair(0, 128, 160)
            i1, i2 = i1.cuda().requires_grad_(True), i2.cuda().requires_grad_(True)
            lo, up = net(i1, i2, iters=iters, test_mode=True)
            (up * gout).sum().backward()
            res[tag] = (up.detach(), i1.grad.clone())
        print(f"iters {iters}: flow cuda-vs-torch {rel(res['cuda'][0], res['torch'][0])}  grad cuda-vs-torch {rel(res['cuda'][1], res['torch'][1])}"
              f"  | rerun: cuda {rel(res['cuda2'][1], res['cuda'][1])} torch {rel(res['torch2'][1], res['torch'][1])}")
