"""torch-CPU restatement of the reference's Python-level operators.  TEST INFRASTRUCTURE ONLY.

Where the reference's algorithm is a composition of PyTorch calls (torch.matmul / F.avg_pool2d /
F.grid_sample / clamp / tanh ...; torch pinned to 1.7.1 in scripts/requirements.txt:2) this module
restates that composition call for call, so it runs on the host cores of any box without the
reference checkout.  It is the checker for the CUDA path at network level, and the body of
`bench.py --impl reference` (kind "port").  Each function cites the lines it follows.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from torch.autograd import Function

from . import ops as cops


# ------------------------------------------------------------------ RAFT / GMA CorrBlock
def bilinear_sampler(img, coords):
    """models/raft/utils/utils.py:57-71."""
    H, W = img.shape[-2:]
    xgrid, ygrid = coords.split([1, 1], dim=-1)
    xgrid = 2 * xgrid / (W - 1) - 1
    ygrid = 2 * ygrid / (H - 1) - 1
    return F.grid_sample(img, torch.cat([xgrid, ygrid], dim=-1), align_corners=True)


class CorrBlock:
    """models/raft/corr.py:12-60 (identical in models/gma/corr.py:15-63)."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        self.num_levels, self.radius = num_levels, radius
        corr = CorrBlock.corr(fmap1, fmap2)
        batch, h1, w1, dim, h2, w2 = corr.shape
        corr = corr.reshape(batch * h1 * w1, dim, h2, w2)
        self.corr_pyramid = [corr]
        for _ in range(num_levels - 1):
            corr = F.avg_pool2d(corr, 2, stride=2)
            self.corr_pyramid.append(corr)

    def __call__(self, coords):
        r = self.radius
        coords = coords.permute(0, 2, 3, 1)
        batch, h1, w1, _ = coords.shape
        out = []
        for i, corr in enumerate(self.corr_pyramid):
            dx = torch.linspace(-r, r, 2 * r + 1)
            dy = torch.linspace(-r, r, 2 * r + 1)
            delta = torch.stack(torch.meshgrid(dy, dx, indexing="ij"), dim=-1).to(coords.device)
            centroid = coords.reshape(batch * h1 * w1, 1, 1, 2) / 2 ** i
            sampled = bilinear_sampler(corr, centroid + delta.view(1, 2 * r + 1, 2 * r + 1, 2))
            out.append(sampled.view(batch, h1, w1, -1))
        return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()

    @staticmethod
    def corr(fmap1, fmap2):
        batch, dim, ht, wd = fmap1.shape
        a = fmap1.view(batch, dim, ht * wd)
        b = fmap2.view(batch, dim, ht * wd)
        corr = torch.matmul(a.transpose(1, 2), b).view(batch, ht, wd, 1, ht, wd)
        return corr / torch.sqrt(torch.tensor(dim).float())


# ------------------------------------------------------------------ spatial correlation sampler
_REF_SCS = None


def _ref_scs_backend():
    """The reference's own CPU extension (oracle/_ref), if it was built; else None."""
    global _REF_SCS
    if _REF_SCS is None:
        try:
            from . import build_ref
            _REF_SCS = build_ref.load() or False
        except Exception:
            _REF_SCS = False
    return _REF_SCS or None


class _ScsFn(Function):
    """Correlation_Module/spatial_correlation_sampler/spatial_correlation_sampler.py:44-91; backend =
    the compiled reference when available (kind "reference"), else the C oracle."""

    @staticmethod
    def forward(ctx, a, b, k, patch, stride, pad, dil, dilp):
        ctx.save_for_backward(a, b)
        ctx.cfg = (k, patch, stride, pad, dil, dilp)
        dev = a.device
        a, b = a.detach().cpu().contiguous(), b.detach().cpu().contiguous()    # CPU-only sampler (PWCNet.py:18-21,48-49)
        be = _ref_scs_backend()
        if be is not None:
            out = be.forward(a, b, k, k, patch, patch, pad, pad, dil, dil, dilp, dilp, stride, stride)
        else:
            out = torch.from_numpy(cops.scs_forward(a.numpy(), b.numpy(), k, patch, stride, pad, dil, dilp))
        return out.to(dev)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        k, patch, stride, pad, dil, dilp = ctx.cfg
        dev = a.device
        a, b, g = a.detach().cpu().contiguous(), b.detach().cpu().contiguous(), g.detach().cpu().contiguous()
        be = _ref_scs_backend()
        if be is not None:
            g1, g2 = be.backward(a, b, g, k, k, patch, patch, pad, pad, dil, dil, dilp, dilp, stride, stride)
        else:
            g1, g2 = cops.scs_backward(a.numpy(), b.numpy(), g.numpy(), k, patch, stride, pad, dil, dilp)
            g1, g2 = torch.from_numpy(g1), torch.from_numpy(g2)
        return g1.to(dev), g2.to(dev), None, None, None, None, None, None


def spatial_correlation_sample(a, b, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1,
                               dilation_patch=1):
    return _ScsFn.apply(a, b, kernel_size, patch_size, stride, padding, dilation, dilation_patch)


def pwc_correlate(input1, input2):
    """models/PWCNet/PWCNet.py:45-58."""
    out = spatial_correlation_sample(input1, input2, kernel_size=1, patch_size=9, stride=1)
    b, ph, pw, h, w = out.size()
    return out.view(b, ph * pw, h, w) / input1.size(1)


def pwc_warp(x, flo):
    """models/PWCNet/PWCNet.py:166-206."""
    B, C, H, W = x.size()
    xx = torch.arange(0, W).view(1, -1).repeat(H, 1)
    yy = torch.arange(0, H).view(-1, 1).repeat(1, W)
    grid = torch.cat((xx.view(1, 1, H, W).repeat(B, 1, 1, 1), yy.view(1, 1, H, W).repeat(B, 1, 1, 1)), 1).float()
    vgrid = grid.to(x.device) + flo
    vx = 2.0 * vgrid[:, 0, :, :] / max(W - 1, 1) - 1.0
    vy = 2.0 * vgrid[:, 1, :, :] / max(H - 1, 1) - 1.0
    vgrid = torch.stack([vx, vy], dim=1).permute(0, 2, 3, 1)
    output = F.grid_sample(x, vgrid, align_corners=False)
    mask = F.grid_sample(torch.ones_like(x), vgrid, align_corners=False)
    mask = (mask >= 0.0001).float()
    return output * mask


# ------------------------------------------------------------------ FlowNet2 operators (C oracle)
class _Fn2CorrFn(Function):
    @staticmethod
    def forward(ctx, a, b, pad, ks, md, s1, s2):
        ctx.save_for_backward(a, b)
        ctx.cfg = (pad, ks, md, s1, s2)
        return torch.from_numpy(cops.fn2corr_forward(a.detach().numpy(), b.detach().numpy(), pad, ks, md, s1, s2))

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g1, g2 = cops.fn2corr_backward(a.detach().numpy(), b.detach().numpy(), g.contiguous().numpy(), *ctx.cfg)
        return torch.from_numpy(g1), torch.from_numpy(g2), None, None, None, None, None


class Correlation(torch.nn.Module):
    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1):
        super().__init__()
        self.cfg = (pad_size, kernel_size, max_displacement, stride1, stride2)

    def forward(self, a, b):
        return _Fn2CorrFn.apply(a, b, *self.cfg)


class _ResampleFn(Function):
    @staticmethod
    def forward(ctx, img, flow):
        ctx.save_for_backward(img, flow)
        return torch.from_numpy(cops.resample2d_forward(img.detach().numpy(), flow.detach().numpy()))

    @staticmethod
    def backward(ctx, g):
        img, flow = ctx.saved_tensors
        gi, gf = cops.resample2d_backward(img.detach().numpy(), flow.detach().numpy(), g.contiguous().numpy())
        return torch.from_numpy(gi), torch.from_numpy(gf)


class Resample2d(torch.nn.Module):
    def __init__(self, kernel_size=1, bilinear=True):
        super().__init__()

    def forward(self, img, flow):
        return _ResampleFn.apply(img.contiguous(), flow.contiguous())


class _ChannelNormFn(Function):
    @staticmethod
    def forward(ctx, x):
        out = torch.from_numpy(cops.channelnorm_forward(x.detach().numpy()))
        ctx.save_for_backward(x, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x, out = ctx.saved_tensors
        return torch.from_numpy(cops.channelnorm_backward(x.detach().numpy(), out.numpy(), g.contiguous().numpy()))


class ChannelNorm(torch.nn.Module):
    def __init__(self, norm_deg=2):
        super().__init__()

    def forward(self, x):
        return _ChannelNormFn.apply(x.contiguous())


# ------------------------------------------------------------------ PCFA objective (torch ops)
def scaled_input(image, delta=None, *, var_change=False, eps_box=0.0, make_unit_input=False):
    """helper_functions/own_models.py:62-85 for one image."""
    if delta is not None:
        image = image + delta.repeat([image.size()[0], 1, 1, 1]) if delta.dim() == 3 or delta.shape[0] == 1 \
            else image + delta
    if var_change:
        image = (1. / 2.) * 1. / (1. - eps_box) * (torch.tanh(image) + (1 - eps_box))
    image = torch.clamp(image, 0., 1.)
    if make_unit_input:
        image = 255. * image
    return image


def extract_deltas(nw_input1, nw_input2, image1, image2, boxconstraint, eps_box=0.):
    """attack_PCFA.py:20-29."""
    if boxconstraint in ['change_of_variables']:
        d1 = (1. / 2.) * 1. / (1. - eps_box) * (torch.tanh(nw_input1) + (1. - eps_box)) - image1
        d2 = (1. / 2.) * 1. / (1. - eps_box) * (torch.tanh(nw_input2) + (1. - eps_box)) - image2
    else:
        d1 = torch.clamp(nw_input1, 0., 1.) - image1
        d2 = torch.clamp(nw_input2, 0., 1.) - image2
    return d1, d2


def extract_deltas_joint(nw_delta, images_max, images_min):
    """attack_PCFA.py:32-37."""
    upper = torch.clamp(nw_delta + images_max, 0., 1.) - images_max
    delta = torch.clamp(upper + images_min, 0., 1.) - images_min
    return delta, delta


def loss_delta_constraint(pred, target, delta1, delta2, device=None, delta_bound=0.001, mu=100., f_type="aee"):
    """helper_functions/losses.py:3-44,76-88,110-126,177-230."""
    if f_type == "aee":
        sq = (pred - target) ** 2
        sim = torch.mean(torch.sum(sq, dim=0 if sq.dim() == 3 else 1).sqrt())
    elif f_type == "mse":
        sim = torch.mean((pred - target) ** 2)
    elif f_type == "cosim":
        sim = 1 - torch.sum(pred * target) / torch.sqrt(torch.sum(pred * pred)) * torch.sqrt(torch.sum(target * target))
    else:
        raise NotImplementedError(f_type)
    numels = torch.numel(delta1) + torch.numel(delta2)
    two_norm = torch.sum(torch.pow(torch.flatten(delta1), 2)) + torch.sum(torch.pow(torch.flatten(delta2), 2))
    pen = torch.max(torch.tensor(0.), two_norm / numels - torch.tensor(delta_bound ** 2))
    return sim + mu * pen
