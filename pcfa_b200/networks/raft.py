"""RAFT (Teed & Deng, ECCV 2020) around the B200 CorrBlock.

Architecturally identical to the network the reference vendors (models/raft/raft.py:24-144,
update.py, extractor.py) and state-dict compatible with its checkpoints (`module.`-prefixed keys of
the DataParallel wrapper are accepted, ownutilities.py:105-107), so a user can load
`raft-sintel.pth` unchanged.  Only the cost-volume operator differs: `corr_block` is injected
(default: pcfa_b200.corr_block.CorrBlock) — tests inject the oracle to obtain reference flows.

The convolutions stay in cuDNN (library code).  On a GPU the glue around them uses this package's fused
kernels: instance norm + ReLU in one op, eval-mode batch norm folded into the frozen
convolution weights, encoders and the RAFT update block in channels-last memory (cuDNN's sm_100 kernels are
NHWC-only), the GRU's element-wise halves fused, its context-feature share hoisted out of the iteration loop.  CPU
tensors take the plain nn-module composition, which is what the oracle-injected parity tests run.
Dead work of the reference's test-mode forward is not executed: the convex-upsampling
mask head and `upsample_flow` run only for the last iteration, because `test_mode=True` returns
nothing else (raft.py:141-142) — outputs and gradients are unchanged.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F


def _norm(kind: str, ch: int) -> nn.Module:
    if kind == "group":
        return nn.GroupNorm(num_groups=ch // 8, num_channels=ch)
    if kind == "batch":
        return nn.BatchNorm2d(ch)
    if kind == "instance":
        return nn.InstanceNorm2d(ch)
    if kind == "none":
        return nn.Sequential()
    raise ValueError(kind)


def _norm_act(norm: nn.Module, x: torch.Tensor, relu: bool) -> torch.Tensor:
    """relu?(norm(x)).  nn.InstanceNorm2d on a GPU runs as one fused pcfa_b200 op (ATen spends 3.1 ms of the 16.8 ms
    RAFT closure on the 15 instance norms + their ReLUs); everything else is the plain module composition."""
    from ..instance_norm import fusable, instance_norm
    if fusable(norm, x):
        return instance_norm(x, eps=norm.eps, relu=relu)
    y = norm(x)
    return F.relu(y) if relu else y


def _bn_folded(conv: nn.Conv2d, bn: nn.BatchNorm2d):
    """Eval-mode BatchNorm is a per-channel affine map: fold it into the preceding convolution's frozen weights
    (W * s, b * s + t with s = gamma / sqrt(var + eps), t = beta - mean * s).  Cached per module; the key notices
    re-allocation and version-counted in-place updates (load_state_dict, copy_, optimiser steps) of any tensor involved —
    not writes through `.data`."""
    ts = (conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
    key = tuple((t.data_ptr(), t._version) for t in ts if t is not None)
    cache = getattr(conv, "_pcfa_bn_fold", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            s_ = torch.rsqrt(bn.running_var + bn.eps)
            if bn.weight is not None:
                s_ = s_ * bn.weight
            t_ = -bn.running_mean * s_
            if bn.bias is not None:
                t_ = t_ + bn.bias
            w = conv.weight * s_.view(-1, 1, 1, 1)
            b = t_ if conv.bias is None else conv.bias * s_ + t_
            if conv.weight.is_contiguous(memory_format=torch.channels_last) and not conv.weight.is_contiguous():
                w = w.contiguous(memory_format=torch.channels_last)
        cache = (key, w, b)
        conv._pcfa_bn_fold = cache
    return cache[1], cache[2]


def _conv_norm_act(conv: nn.Conv2d, norm: nn.Module, x: torch.Tensor, relu: bool) -> torch.Tensor:
    """relu?(norm(conv(x))) with the two GPU fast paths of the encoders: instance norm + ReLU as one fused pcfa_b200 op,
    frozen eval-mode batch norm folded into the convolution."""
    frozen = not any(p.requires_grad for p in conv.parameters())
    if (x.is_cuda and isinstance(norm, nn.BatchNorm2d) and not norm.training and norm.track_running_stats and frozen
            and conv.padding_mode == "zeros"):
        w, b = _bn_folded(conv, norm)
        from ..conv_ops import conv_act                     # bias + ReLU as the convolution's fused epilogue (csrc/bias_act.cu)
        return conv_act(conv, x, relu, w, b, "_pcfa_bn_fold16")
    if (x.is_cuda and frozen and isinstance(norm, nn.InstanceNorm2d) and not norm.affine and not norm.track_running_stats
            and conv.padding_mode == "zeros" and conv.bias is not None):
        # a per-channel constant is removed by the instance norm that follows: skip the bias add (a broadcasting ATen
        # launch over up to 58 MB) — (conv + b) - mean(conv + b) == conv - mean(conv)
        from .amp import amp_half_active, half_params
        w = half_params(conv, conv.weight, None, "_pcfa_w16_nobias")[0] if amp_half_active(x) else conv.weight
        return _norm_act(norm, F.conv2d(x, w, None, conv.stride, conv.padding, conv.dilation, conv.groups), relu)
    return _norm_act(norm, conv(x), relu)


class ResidualBlock(nn.Module):
    def __init__(self, cin, cout, norm_fn="group", stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.norm1 = _norm(norm_fn, cout)
        self.norm2 = _norm(norm_fn, cout)
        self.downsample = None
        if stride != 1:
            self.norm3 = _norm(norm_fn, cout)       # registered twice (norm3 / downsample.1) like the reference
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride), self.norm3)

    def forward(self, x):
        # the previous block's tail hands over two autograd handles of its result (conv_ops._AddReLUTwin): one for the
        # convolution path, one for the skip branch, so that their gradients meet inside that tail's ReLU mask
        skip = getattr(x, "_pcfa_twin", None)
        skip = x if skip is None else skip
        y = _conv_norm_act(self.conv1, self.norm1, x, True)
        y = _conv_norm_act(self.conv2, self.norm2, y, True)
        if self.downsample is not None:
            skip = _conv_norm_act(self.downsample[0], self.downsample[1], skip, False)
        from ..conv_ops import add_relu                     # relu(skip + y) in one launch on the GPU
        return add_relu(skip, y, twin=True)


class BottleneckBlock(nn.Module):
    def __init__(self, cin, cout, norm_fn="group", stride=1):
        super().__init__()
        q = cout // 4
        self.conv1 = nn.Conv2d(cin, q, 1)
        self.conv2 = nn.Conv2d(q, q, 3, padding=1, stride=stride)
        self.conv3 = nn.Conv2d(q, cout, 1)
        self.relu = nn.ReLU(inplace=True)
        if norm_fn == "group":
            g = cout // 8
            self.norm1, self.norm2 = nn.GroupNorm(g, q), nn.GroupNorm(g, q)
            self.norm3 = nn.GroupNorm(g, cout)
        else:
            self.norm1, self.norm2, self.norm3 = _norm(norm_fn, q), _norm(norm_fn, q), _norm(norm_fn, cout)
        self.downsample = None
        if stride != 1:
            self.norm4 = nn.GroupNorm(cout // 8, cout) if norm_fn == "group" else _norm(norm_fn, cout)
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride), self.norm4)

    def forward(self, x):
        y = _conv_norm_act(self.conv1, self.norm1, x, True)
        y = _conv_norm_act(self.conv2, self.norm2, y, True)
        y = _conv_norm_act(self.conv3, self.norm3, y, True)
        if self.downsample is not None:
            x = _conv_norm_act(self.downsample[0], self.downsample[1], x, False)
        from ..conv_ops import add_relu                     # relu(x + y) in one launch on the GPU
        return add_relu(x, y)


def _init_encoder(mod: nn.Module):
    for m in mod.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d, nn.GroupNorm)):
            if m.weight is not None:
                nn.init.constant_(m.weight, 1)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)


class _Encoder(nn.Module):
    """Shared trunk logic of BasicEncoder / SmallEncoder: stem, three stages, 1x1 head."""
    block = ResidualBlock
    widths = (64, 64, 96, 128)

    def __init__(self, output_dim=128, norm_fn="batch", dropout=0.0):
        super().__init__()
        self.norm_fn = norm_fn
        stem, w1, w2, w3 = self.widths
        self.norm1 = nn.GroupNorm(8, stem) if norm_fn == "group" else _norm(norm_fn, stem)
        self.conv1 = nn.Conv2d(3, stem, 7, stride=2, padding=3)
        self.relu1 = nn.ReLU(inplace=True)
        self.in_planes = stem
        self.layer1 = self._stage(w1, 1)
        self.layer2 = self._stage(w2, 2)
        self.layer3 = self._stage(w3, 2)
        self.conv2 = nn.Conv2d(w3, output_dim, 1)
        self.dropout = nn.Dropout2d(p=dropout) if dropout > 0 else None
        _init_encoder(self)

    def _stage(self, dim, stride):
        blocks = nn.Sequential(self.block(self.in_planes, dim, self.norm_fn, stride=stride),
                               self.block(dim, dim, self.norm_fn, stride=1))
        self.in_planes = dim
        return blocks

    def forward(self, x):
        pair = isinstance(x, (tuple, list))
        if pair:
            nb = x[0].shape[0]
            x = torch.cat(x, dim=0)
        cl = x.is_cuda and getattr(self, "channels_last", False)
        if cl:                                              # NHWC through the whole encoder: cuDNN's sm_100 kernels are
            x = x.contiguous(memory_format=torch.channels_last)      # NHWC-only and otherwise convert around every conv
        x = _conv_norm_act(self.conv1, self.norm1, x, True)
        x = self.conv2(self.layer3(self.layer2(self.layer1(x))))
        if cl:
            x = x.contiguous()
        if self.training and self.dropout is not None:
            x = self.dropout(x)
        return torch.split(x, [nb, nb], dim=0) if pair else x


class BasicEncoder(_Encoder):
    block = ResidualBlock
    widths = (64, 64, 96, 128)


class SmallEncoder(_Encoder):
    block = BottleneckBlock
    widths = (32, 32, 64, 96)


class FlowHead(nn.Module):
    def __init__(self, input_dim=128, hidden_dim=256):
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, 2, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        from ..conv_ops import conv_act, padded_out_channels
        y = conv_act(self.conv1, x, True)
        frozen = not any(p.requires_grad for p in self.conv2.parameters())
        if y.is_cuda and frozen and y.is_contiguous(memory_format=torch.channels_last) and not y.is_contiguous() \
                and os.environ.get("PCFA_PAD_CHANNELS", "1") != "0":
            w, b = padded_out_channels(self.conv2, 8)            # 2 -> 8 output channels (zero filters): see padded_out_channels
            return conv_act(self.conv2, y, False, w, b, "_pcfa_pad16")[:, :self.conv2.out_channels]
        return conv_act(self.conv2, y, False)


def _zr_weights(cz: nn.Conv2d, cr: nn.Conv2d):
    """convz and convr read the same input: run them as one convolution with concatenated output channels (frozen
    weights; cached, the key notices in-place updates and re-allocation)."""
    ts = (cz.weight, cz.bias, cr.weight, cr.bias)
    key = tuple((t.data_ptr(), t._version) for t in ts)
    cache = getattr(cz, "_pcfa_zr", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            w = torch.cat([cz.weight, cr.weight], dim=0)
            cl = cz.weight.is_contiguous(memory_format=torch.channels_last) and not cz.weight.is_contiguous()
            w = w.contiguous(memory_format=torch.channels_last) if cl else w.contiguous()
            cache = (key, w, torch.cat([cz.bias, cr.bias], dim=0).contiguous())
        cz._pcfa_zr = cache
    return cache[1], cache[2]


def _gru_step(h, x, cz, cr, cq, cl=False):
    from .. import gru_ops
    frozen = not any(p.requires_grad for m in (cz, cr) for p in m.parameters())
    if gru_ops.usable(h, x) and frozen and cz.bias is not None and cr.bias is not None:
        # GPU path: one convolution for both gates, two fused element-wise launches (pcfa_b200/csrc/gru.cu)
        hx = gru_ops.cat_channels([h, x], cl)
        w, b = _zr_weights(cz, cr)
        z, rh = gru_ops.gru_gates(F.conv2d(hx, w, b, cz.stride, cz.padding), h)
        return gru_ops.gru_blend(z, cq(gru_ops.cat_channels([rh, x], cl)), h)
    hx = torch.cat([h, x], dim=1)
    z = torch.sigmoid(cz(hx))
    r = torch.sigmoid(cr(hx))
    q = torch.tanh(cq(torch.cat([r * h, x], dim=1)))
    return (1 - z) * h + z * q


class ConvGRU(nn.Module):
    def __init__(self, hidden_dim=128, input_dim=192 + 128):
        super().__init__()
        c = hidden_dim + input_dim
        self.convz = nn.Conv2d(c, hidden_dim, 3, padding=1)
        self.convr = nn.Conv2d(c, hidden_dim, 3, padding=1)
        self.convq = nn.Conv2d(c, hidden_dim, 3, padding=1)

    def forward(self, h, x):
        return _gru_step(h, x, self.convz, self.convr, self.convq)


class SepConvGRU(nn.Module):
    def __init__(self, hidden_dim=128, input_dim=192 + 128):
        super().__init__()
        c = hidden_dim + input_dim
        for tag, k, p in (("1", (1, 5), (0, 2)), ("2", (5, 1), (2, 0))):
            for gate in "zrq":
                setattr(self, f"conv{gate}{tag}", nn.Conv2d(c, hidden_dim, k, padding=p))

    def forward(self, h, x, cl=False):
        h = _gru_step(h, x, self.convz1, self.convr1, self.convq1, cl)      # horizontal
        return _gru_step(h, x, self.convz2, self.convr2, self.convq2, cl)   # vertical

    # ---- NHWC fast path: x = [inp | motion] and inp is the same in every iteration, so its share of the four
    # convolutions is computed once per forward pass ("hoisted") and enters the fused element-wise kernels as an addend
    def _split_weights(self, ci):
        convs = [self.convz1, self.convr1, self.convq1, self.convz2, self.convr2, self.convq2]
        key = tuple((t.data_ptr(), t._version) for c in convs for t in (c.weight, c.bias)) + (ci,)
        cache = getattr(self, "_pcfa_split", None)
        if cache is None or cache[0] != key:
            ch = self.convz1.out_channels
            out = []
            with torch.no_grad():
                for cz, cr, cq in ((self.convz1, self.convr1, self.convq1), (self.convz2, self.convr2, self.convq2)):
                    step = []
                    for w, b in ((torch.cat([cz.weight, cr.weight], 0), torch.cat([cz.bias, cr.bias], 0)), (cq.weight, cq.bias)):
                        w_hm = torch.cat([w[:, :ch], w[:, ch + ci:]], 1).contiguous(memory_format=torch.channels_last)
                        w_in = w[:, ch:ch + ci].contiguous(memory_format=torch.channels_last)
                        step.append((w_hm, w_in, b.contiguous()))
                    out.append((step, cz.padding))
            cache = (key, out)
            self._pcfa_split = cache
        return cache[1]

    def hoisted(self, inp):
        """[(W_hm_zr, P_zr, W_hm_q, P_q, padding)] for the horizontal and the vertical step; P = conv(inp, W_inp) + bias."""
        res = []
        for (zr, q), pad in self._split_weights(inp.shape[1]):
            res.append((zr[0], F.conv2d(inp, zr[1], zr[2], 1, pad), q[0], F.conv2d(inp, q[1], q[2], 1, pad), pad))
        return res

    def forward_x(self, h, motion, hoist):
        from ..gru_ops import cat_channels, gru_blend_x, gru_gates_x
        (wzr1, pzr1, wq1, pq1, pad1), (wzr2, pzr2, wq2, pq2, pad2) = hoist
        hm = cat_channels([h, motion], True)
        z, rhm = gru_gates_x(F.conv2d(hm, wzr1, None, 1, pad1), pzr1, h, motion)
        h, hm = gru_blend_x(z, F.conv2d(rhm, wq1, None, 1, pad1), pq1, h, motion, True)
        z, rhm = gru_gates_x(F.conv2d(hm, wzr2, None, 1, pad2), pzr2, h, motion)
        return gru_blend_x(z, F.conv2d(rhm, wq2, None, 1, pad2), pq2, h, motion, False)


class BasicMotionEncoder(nn.Module):
    def __init__(self, corr_levels, corr_radius):
        super().__init__()
        cor_planes = corr_levels * (2 * corr_radius + 1) ** 2
        self.convc1 = nn.Conv2d(cor_planes, 256, 1)
        self.convc2 = nn.Conv2d(256, 192, 3, padding=1)
        self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(64 + 192, 128 - 2, 3, padding=1)

    def forward(self, flow, corr, cl=False):
        from ..gru_ops import cat_channels
        from ..conv_ops import conv_act
        cor = conv_act(self.convc2, conv_act(self.convc1, corr, True), True)
        if flow.shape[1] > self.convf1.in_channels:         # zero-padded flow (conv_ops.flow_step): 2 -> 8 input channels turn the
            from ..conv_ops import padded_in_channels       # 7x7 convolution into a tensor-core implicit GEMM (13.5 -> ~4 us)
            f1 = conv_act(self.convf1, flow, True, padded_in_channels(self.convf1, flow.shape[1]), self.convf1.bias, "_pcfa_padin16")
            flow = flow[:, :self.convf1.in_channels]
        else:
            f1 = conv_act(self.convf1, flow, True)
        flo = conv_act(self.convf2, f1, True)
        x = cat_channels([cor, flo], cl)
        frozen = not any(p.requires_grad for p in self.conv.parameters())
        if cl and frozen and x.is_cuda and not flow.requires_grad and (self.conv.out_channels + flow.shape[1]) % 8 == 0 \
                and os.environ.get("PCFA_PAD_CHANNELS", "1") != "0":
            # 126 -> 128 output channels with two zero filters, the flow written over them: no cuDNN channel padding
            # around the convolution and its data gradient, and torch.cat([out, flow]) costs a 56 KB strided write
            from ..conv_ops import padded_out_channels
            w, b = padded_out_channels(self.conv, 8)
            if w.shape[0] == self.conv.out_channels + flow.shape[1]:
                return conv_act(self.conv, x, True, w, b, "_pcfa_pad16", tail=flow)
        out = conv_act(self.conv, x, True)
        return cat_channels([out, flow], cl)


class SmallMotionEncoder(nn.Module):
    def __init__(self, corr_levels, corr_radius):
        super().__init__()
        cor_planes = corr_levels * (2 * corr_radius + 1) ** 2
        self.convc1 = nn.Conv2d(cor_planes, 96, 1)
        self.convf1 = nn.Conv2d(2, 64, 7, padding=3)
        self.convf2 = nn.Conv2d(64, 32, 3, padding=1)
        self.conv = nn.Conv2d(128, 80, 3, padding=1)

    def forward(self, flow, corr):
        cor = F.relu(self.convc1(corr))
        flo = F.relu(self.convf2(F.relu(self.convf1(flow))))
        out = F.relu(self.conv(torch.cat([cor, flo], dim=1)))
        return torch.cat([out, flow], dim=1)


class BasicUpdateBlock(nn.Module):
    def __init__(self, corr_levels, corr_radius, hidden_dim=128):
        super().__init__()
        self.encoder = BasicMotionEncoder(corr_levels, corr_radius)
        self.gru = SepConvGRU(hidden_dim=hidden_dim, input_dim=128 + hidden_dim)
        self.flow_head = FlowHead(hidden_dim, hidden_dim=256)
        self.mask = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(inplace=True),
                                  nn.Conv2d(256, 64 * 9, 1))

    def forward(self, net, inp, corr, flow, want_mask=True, cl=False, hoist=None, raw_mask=False, step_sources=None, last=False):
        """cl=True: every tensor is torch.channels_last (cuDNN's sm_100 kernels are NHWC-only; with NCHW activations
        it converts around every convolution: 3 ms of the 11.8 ms RAFT closure)."""
        from ..gru_ops import cat_channels
        motion = self.encoder(flow, corr, cl)
        if step_sources is not None:                      # one autograd node for the whole GRU step (gru_ops.gru_step_x)
            from ..gru_ops import gru_step_x
            net = gru_step_x(net, motion, step_sources, last)
        elif hoist is not None:
            net = self.gru.forward_x(net, motion, hoist)
        else:
            net = self.gru(net, cat_channels([inp, motion], cl), cl)
        delta_flow = self.flow_head(net)
        # .25 "to balance gradients" (update.py:135); raw_mask=True leaves it to the fused up-sampling kernel's mask_scale
        mask = None
        if want_mask:
            from ..conv_ops import conv_act
            mask = conv_act(self.mask[2], conv_act(self.mask[0], net, True), False)
            if not raw_mask:
                mask = 0.25 * mask
        return net, mask, delta_flow


class SmallUpdateBlock(nn.Module):
    def __init__(self, corr_levels, corr_radius, hidden_dim=96):
        super().__init__()
        self.encoder = SmallMotionEncoder(corr_levels, corr_radius)
        self.gru = ConvGRU(hidden_dim=hidden_dim, input_dim=82 + 64)
        self.flow_head = FlowHead(hidden_dim, hidden_dim=128)

    def forward(self, net, inp, corr, flow, want_mask=True):
        motion = self.encoder(flow, corr)
        net = self.gru(net, torch.cat([inp, motion], dim=1))
        return net, None, self.flow_head(net)


def coords_grid(batch, ht, wd, device):
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


def upflow8(flow, mode="bilinear"):
    new_size = (8 * flow.shape[2], 8 * flow.shape[3])
    return 8 * F.interpolate(flow, size=new_size, mode=mode, align_corners=True)


def convex_upsample(flow, mask):
    """[N,2,H,W] → [N,2,8H,8W] by a learned convex combination of the 3x3 neighbourhood (raft.py:72-83)."""
    N, _, H, W = flow.shape
    flow, mask = flow.contiguous(), mask.contiguous()      # (channels-last update block: back to NCHW for the views)
    mask = torch.softmax(mask.view(N, 1, 9, 8, 8, H, W), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(N, 2, 9, 1, 1, H, W)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(N, 2, 8 * H, 8 * W)


DEFAULT_CONFIG = {"small": False, "mixed_precision": False, "epsilon": 1e-8}


class RAFT(nn.Module):
    def __init__(self, args=None, corr_block=None):
        super().__init__()
        args = dict(DEFAULT_CONFIG if args is None else (vars(args) if not isinstance(args, dict) else args))
        self.args = args
        small = bool(args.get("small", False))
        if small:
            self.hidden_dim, self.context_dim = 96, 64
            args["corr_levels"], args["corr_radius"] = 4, 3
        else:
            self.hidden_dim, self.context_dim = 128, 128
            args["corr_levels"], args["corr_radius"] = 4, 4
        args.setdefault("dropout", 0)
        args.setdefault("mixed_precision", False)
        hd, cd = self.hidden_dim, self.context_dim
        if small:
            self.fnet = SmallEncoder(output_dim=128, norm_fn="instance", dropout=args["dropout"])
            self.cnet = SmallEncoder(output_dim=hd + cd, norm_fn="none", dropout=args["dropout"])
            self.update_block = SmallUpdateBlock(4, 3, hidden_dim=hd)
        else:
            self.fnet = BasicEncoder(output_dim=256, norm_fn="instance", dropout=args["dropout"])
            self.cnet = BasicEncoder(output_dim=hd + cd, norm_fn="batch", dropout=args["dropout"])
            self.update_block = BasicUpdateBlock(4, 4, hidden_dim=hd)
        if corr_block is None:
            from ..corr_block import CorrBlock as corr_block
        self.corr_block = corr_block

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        return super().load_state_dict(sd, strict=strict, **kw)

    def forward(self, image1, image2, iters=12, flow_init=None, upsample=True, test_mode=False):
        image1 = (2 * (image1 / 255.0) - 1.0).contiguous()
        image2 = (2 * (image2 / 255.0) - 1.0).contiguous()
        dev_type = image1.device.type
        # torch.cuda.amp.autocast of the reference is a no-op off-GPU (raft.py:12)
        amp = bool(self.args["mixed_precision"]) and dev_type == "cuda"
        with torch.autocast(dev_type, enabled=amp):
            fmap1, fmap2 = self.fnet([image1, image2])
        corr_fn = self.corr_block(fmap1.float(), fmap2.float(), radius=self.args["corr_radius"])
        with torch.autocast(dev_type, enabled=amp):
            net, inp = torch.split(self.cnet(image1), [self.hidden_dim, self.context_dim], dim=1)
            net, inp = torch.tanh(net), torch.relu(inp)
        N, _, H, W = image1.shape
        coords0 = coords_grid(N, H // 8, W // 8, image1.device)
        coords1 = coords0.clone()
        if flow_init is not None:
            coords1 = coords1 + flow_init
        predictions = []
        flow_up = None
        # NHWC update block (GPU, own CorrBlock, fp32): the lookup emits channels-last features directly
        from ..corr_block import CorrBlock as _OwnCorrBlock
        cl = (bool(getattr(self.update_block, "channels_last", False)) and dev_type == "cuda" and not amp
              and isinstance(corr_fn, _OwnCorrBlock) and isinstance(self.update_block, BasicUpdateBlock))
        hoist = step_sources = None
        if cl:
            net = net.contiguous(memory_format=torch.channels_last)
            inp = inp.contiguous(memory_format=torch.channels_last)
            frozen = not any(p.requires_grad for p in self.update_block.gru.parameters())
            if frozen and net.shape[1] % 4 == 0 and inp.shape[1] % 4 == 0:
                hoist = self.update_block.gru.hoisted(inp)
                if torch.is_grad_enabled() and os.environ.get("PCFA_GRU_STEP", "1") != "0":
                    from ..gru_ops import hoist_sources
                    step_sources = hoist_sources(hoist)
        flow_cl = None
        for itr in range(iters):
            coords1 = coords1.detach()
            corr = corr_fn(coords1, channels_last=True) if cl else corr_fn(coords1)
            need_up = (not test_mode) or itr == iters - 1
            with torch.autocast(dev_type, enabled=amp):
                if cl:
                    from ..conv_ops import flow_step, padded_flow
                    fch = 8 if os.environ.get("PCFA_PAD_CHANNELS", "1") != "0" else 2
                    if flow_cl is None:
                        flow_cl = padded_flow(coords1 - coords0, fch)
                    net, up_mask, delta_flow = self.update_block(net, inp, corr, flow_cl,
                                                                 want_mask=need_up, cl=True, hoist=hoist, raw_mask=True,
                                                                 step_sources=step_sources, last=itr == iters - 1)
                    # coords1 += delta_flow and the next iteration's channels-last flow in one launch
                    coords1, flow_cl = flow_step(coords1, coords0, delta_flow, fch)
                else:
                    flow = coords1 - coords0
                    net, up_mask, delta_flow = self.update_block(net, inp, corr, flow, want_mask=need_up)
                    coords1 = coords1 + delta_flow
            if need_up:
                if up_mask is None:
                    flow_up = upflow8(coords1 - coords0)
                elif cl:                             # fused kernel on the raw channels-last mask (csrc/upsample.cu)
                    from ..upsample import convex_upsample as _fused_upsample
                    flow_up = _fused_upsample(coords1 - coords0, up_mask, 0.25)
                else:
                    flow_up = convex_upsample(coords1 - coords0, up_mask)
                predictions.append(flow_up)
        if test_mode:
            return coords1 - coords0, flow_up
        return predictions
