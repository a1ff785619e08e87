"""The 15 instance-norm (+ReLU) calls of RAFT's feature encoder at the BASELINE shape (2 images, channels_last),
forward and backward — the short command the ncu capture of the instnorm kernels is taken from."""
import sys, torch
sys.path.insert(0, '.')
from pcfa_b200.instance_norm import instance_norm
shapes = [(2, 64, 220, 512)] * 5 + [(2, 96, 110, 256)] * 5 + [(2, 128, 55, 128)] * 5
only = int(sys.argv[1]) if len(sys.argv) > 1 else None
for it in range(2):
    for i, s in enumerate(shapes):
        if only is not None and i % 5 != only:
            continue
        x = torch.randn(s, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y = instance_norm(x, relu=True)
        y.backward(torch.randn_like(y))
torch.cuda.synchronize()
print("done")
