"""Deterministic, name-keyed synthetic weights (there is no network access for checkpoints).

Every tensor of a state_dict is filled from a generator seeded by crc32(key) ^ seed, so two
implementations of the same architecture (this package's and the reference's) receive bit-identical
weights as long as their state_dict keys and shapes agree — which is also the checkpoint-compatibility
contract.  Scales follow He initialisation so activations stay O(1) through the stack; `gain` < 1 damps
the stack (gain 0.5 gives RAFT flows of a few pixels, like a trained network, instead of ~100 px).
"""
from __future__ import annotations

import math
import zlib

import torch


def deterministic_state_(model: torch.nn.Module, seed: int = 0, strip_prefix: str = "", gain: float = 1.0) -> torch.nn.Module:
    sd = model.state_dict()
    with torch.no_grad():
        for key in sorted(sd.keys()):
            t = sd[key]
            if not t.is_floating_point():
                continue
            name = key[len(strip_prefix):] if strip_prefix and key.startswith(strip_prefix) else key
            g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ seed) & 0x7FFFFFFF)
            if name.endswith("running_var"):
                v = torch.rand(t.shape, generator=g) + 0.5
            elif name.endswith("running_mean"):
                v = 0.1 * torch.randn(t.shape, generator=g)
            elif t.dim() >= 2:
                fan_in = t[0].numel()
                v = torch.randn(t.shape, generator=g) * (gain * math.sqrt(2.0 / fan_in))
            elif name.endswith("weight"):
                v = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
            else:
                v = 0.05 * torch.randn(t.shape, generator=g)
            t.copy_(v.to(t.dtype))
    return model


def synthetic_pair(idx: int, H: int, W: int, batch: int = 1):
    """Synthetic image pair in [0,255]: image2 = roll(image1, (3,5)) + N(0, 2^2)."""
    g = torch.Generator().manual_seed(1234 + idx)
    img1 = torch.rand(batch, 3, H, W, generator=g) * 255.0
    # low-pass so that the pair has structure a flow network can lock on to
    img1 = torch.nn.functional.avg_pool2d(img1, 5, stride=1, padding=2, count_include_pad=False)
    img1 = (img1 - img1.min()) / (img1.max() - img1.min()) * 255.0
    img2 = torch.roll(img1, shifts=(3, 5), dims=(2, 3)) + 2.0 * torch.randn(batch, 3, H, W, generator=g)
    return img1.clamp(0, 255).contiguous(), img2.clamp(0, 255).contiguous()
