"""GMA (Jiang et al., ICCV 2021) around the B200 CorrBlock.

Same architecture and state-dict keys as the reference's vendored network (models/gma/network.py:
21-129, gma.py:34-115, update.py:112-139); the cost volume is this package's CorrBlock.  The
reference evaluates GMA with 6 iterations and fp16 autocast on CUDA (ownutilities.py:327,
models/_config/gma_config.json:5); both are kept.  As in networks/raft.py the mask head and convex
upsampling only run for the prediction that test_mode returns.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
from torch import einsum

from .raft import BasicEncoder, BasicMotionEncoder, FlowHead, SepConvGRU, convex_upsample, coords_grid

DEFAULT_CONFIG = {"epsilon": 1e-8, "num_heads": 1, "small": False, "mixed_precision": True,
                  "position_only": False, "position_and_content": False}


class RelPosEmb(nn.Module):
    def __init__(self, max_pos_size, dim_head):
        super().__init__()
        self.rel_height = nn.Embedding(2 * max_pos_size - 1, dim_head)
        self.rel_width = nn.Embedding(2 * max_pos_size - 1, dim_head)
        deltas = torch.arange(max_pos_size).view(1, -1) - torch.arange(max_pos_size).view(-1, 1)
        self.register_buffer('rel_ind', deltas + max_pos_size - 1)

    def forward(self, q):
        _, _, h, w, _ = q.shape
        he = self.rel_height(self.rel_ind[:h, :h].reshape(-1)).view(h, h, 1, -1)     # (x u) d -> x u () d
        we = self.rel_width(self.rel_ind[:w, :w].reshape(-1)).view(w, 1, w, -1)      # (y v) d -> y () v d
        return einsum('b h x y d, x u v d -> b h x y u v', q, he) + einsum('b h x y d, y u v d -> b h x y u v', q, we)


class Attention(nn.Module):
    def __init__(self, *, position_only=False, position_and_content=False, dim, max_pos_size=100, heads=4, dim_head=128):
        super().__init__()
        self.position_only, self.position_and_content = position_only, position_and_content
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_qk = nn.Conv2d(dim, heads * dim_head * 2, 1, bias=False)
        self.pos_emb = RelPosEmb(max_pos_size, dim_head)

    def forward(self, fmap):
        b, _, h, w = fmap.shape
        q, k = self.to_qk(fmap).chunk(2, dim=1)
        q = q.view(b, self.heads, -1, h, w).permute(0, 1, 3, 4, 2) * self.scale     # b h x y d
        k = k.view(b, self.heads, -1, h, w).permute(0, 1, 3, 4, 2)
        if self.position_only:
            sim = self.pos_emb(q)
        else:
            sim = einsum('b h x y d, b h u v d -> b h x y u v', q, k)
            if self.position_and_content:
                sim = sim + self.pos_emb(q)
        sim = sim.reshape(b, self.heads, h * w, h * w)
        from ..attention import softmax_rows_f16, softmax_rows_supported
        if softmax_rows_supported(sim):
            # fp16 autocast (the shipped configuration): one fused pass, fp16 in / fp16 out, fp32 arithmetic — the
            # values the aggregation GEMMs consume in the reference (fp32 softmax, cast to fp16 by every einsum)
            return softmax_rows_f16(sim)
        return sim.softmax(dim=-1)


class _AttnSource(torch.autograd.Function):
    """Identity on the attention matrix.  Every iteration's aggregation consumes the returned tensor and returns no gradient
    for it; it only files its (g_k, v_k) pair in the shared holder.  This node runs after all of them and computes
    dL/dattn = sum_k g_k v_k^T as ONE GEMM with K = iterations x 128 — instead of six K = 128 GEMMs that each write a 99 MB
    fp16 matrix plus five adds of those matrices (0.45 ms -> 0.1 ms per closure)."""

    @staticmethod
    def forward(ctx, attn, holder):
        ctx.holder = holder
        ctx.set_materialize_grads(False)
        return attn.view_as(attn)

    @staticmethod
    def backward(ctx, g):
        gs, vs = ctx.holder.pop("g", None), ctx.holder.pop("v", None)
        if not gs:
            return g, None
        G = gs[0] if len(gs) == 1 else torch.cat(gs, dim=2)             # [b, n, K]
        V = vs[0] if len(vs) == 1 else torch.cat(vs, dim=2)
        acc = torch.bmm(G, V.transpose(1, 2)).view(ctx.holder["shape"])
        return (acc if g is None else acc + g), None


class _AttnBmm(torch.autograd.Function):
    """attn · v for one iteration; dL/dv here, the dL/dattn share deferred to _AttnSource."""

    @staticmethod
    def forward(ctx, a2, vm, holder):
        ctx.save_for_backward(a2, vm)
        ctx.holder = holder
        return torch.bmm(a2, vm)

    @staticmethod
    def backward(ctx, g):
        a2, vm = ctx.saved_tensors
        gvm = torch.bmm(a2.transpose(1, 2), g) if ctx.needs_input_grad[1] else None
        ctx.holder.setdefault("g", []).append(g)
        ctx.holder.setdefault("v", []).append(vm)
        return None, gvm, None


class Aggregate(nn.Module):
    def __init__(self, dim, heads=4, dim_head=128):
        super().__init__()
        self.heads = heads
        self.scale = dim_head ** -0.5
        inner = heads * dim_head
        self.to_v = nn.Conv2d(dim, inner, 1, bias=False)
        self.gamma = nn.Parameter(torch.zeros(1))
        self.project = nn.Conv2d(inner, dim, 1, bias=False) if dim != inner else None

    def forward(self, attn, fmap):
        b, _, h, w = fmap.shape
        v = self.to_v(fmap)
        if self.heads == 1 and v.is_cuda and v.is_contiguous(memory_format=torch.channels_last) and not v.is_contiguous():
            # channels-last memory IS 'b (x y) d': the aggregation is one batched GEMM on views, no copies either side
            vm = v.permute(0, 2, 3, 1).reshape(b, h * w, -1)
            a2 = attn.reshape(b, h * w, h * w)
            holder = getattr(attn, "_pcfa_holder", None)
            if holder is not None and a2.dtype == vm.dtype and torch.is_grad_enabled():
                out = _AttnBmm.apply(a2, vm, holder)
            else:
                out = torch.bmm(a2.to(vm.dtype) if a2.dtype != vm.dtype else a2, vm)
            out = out.view(b, h, w, -1).permute(0, 3, 1, 2)                         # channels-last [b, d, h, w]
            if self.project is not None:
                out = self.project(out)
            return fmap + self.gamma * out
        v = v.reshape(b, self.heads, -1, h * w).transpose(2, 3)                     # b h (x y) d
        out = einsum('b h i j, b h j d -> b h i d', attn, v)
        out = out.transpose(2, 3).reshape(b, -1, h, w)                              # b (h d) x y
        if self.project is not None:
            out = self.project(out)
        return fmap + self.gamma * out


class GMAUpdateBlock(nn.Module):
    def __init__(self, corr_levels, corr_radius, num_heads, hidden_dim=128):
        super().__init__()
        self.encoder = BasicMotionEncoder(corr_levels, corr_radius)
        self.gru = SepConvGRU(hidden_dim=hidden_dim, input_dim=128 + hidden_dim + hidden_dim)
        self.flow_head = FlowHead(hidden_dim, hidden_dim=256)
        self.mask = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(256, 64 * 9, 1))
        self.aggregator = Aggregate(dim=128, dim_head=128, heads=num_heads)

    def forward(self, net, inp, corr, flow, attention, want_mask=True, raw_mask=False, step_sources=None, last=False):
        """With channels-last inputs (and channels-last weights, adapter.build_network) every tensor stays NHWC: cuDNN's
        sm_100 kernels are NHWC-only and otherwise convert around each of the ~17 convolutions per iteration.
        step_sources (gru_ops.hoist_sources): the whole SepConvGRU step as one autograd node with the context features'
        share of its convolutions hoisted out of the iteration loop (fp16 kernels under autocast, csrc/gru_half.cu)."""
        # channels-last: the encoder's last convolution gets two zero filters and the flow written over them (no 126-channel
        # tensor for cuDNN to pad, no concatenation), as in RAFT's update block
        motion = self.encoder(flow, corr, bool(getattr(self, "channels_last", False)) and flow.is_cuda)
        motion_global = self.aggregator(attention, motion)
        if step_sources is not None:
            from ..gru_ops import gru_step_x
            dt = step_sources[0][1].dtype
            net = gru_step_x(net, torch.cat([motion.to(dt), motion_global.to(dt)], dim=1), step_sources, last)
        else:
            net = self.gru(net, torch.cat([inp, motion, motion_global], dim=1))
        delta_flow = self.flow_head(net)
        mask = None
        if want_mask:
            from ..conv_ops import conv_act
            mask = conv_act(self.mask[2], conv_act(self.mask[0], net, True), False)
            if not raw_mask:
                mask = 0.25 * mask
        return net, mask, delta_flow


class RAFTGMA(nn.Module):
    def __init__(self, args=None, corr_block=None):
        super().__init__()
        cfg = dict(DEFAULT_CONFIG)
        if args is not None:
            cfg.update(vars(args) if not isinstance(args, dict) else args)
        cfg.setdefault("dropout", 0)
        cfg["corr_levels"], cfg["corr_radius"] = 4, 4
        self.args = cfg
        self.hidden_dim = self.context_dim = 128
        self.fnet = BasicEncoder(output_dim=256, norm_fn='instance', dropout=cfg["dropout"])
        self.cnet = BasicEncoder(output_dim=256, norm_fn='batch', dropout=cfg["dropout"])
        self.update_block = GMAUpdateBlock(4, 4, cfg["num_heads"], hidden_dim=128)
        self.att = Attention(position_only=cfg["position_only"], position_and_content=cfg["position_and_content"],
                             dim=128, heads=cfg["num_heads"], max_pos_size=160, dim_head=128)
        if corr_block is None:
            from ..corr_block import CorrBlock as corr_block
        self.corr_block = corr_block

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        return super().load_state_dict(sd, strict=strict, **kw)

    def forward(self, image1, image2, iters=12, flow_init=None, upsample=True, test_mode=False):
        image1 = (2 * (image1 / 255.0) - 1.0).contiguous()
        image2 = (2 * (image2 / 255.0) - 1.0).contiguous()
        dev_type = image1.device.type
        amp = bool(self.args["mixed_precision"]) and dev_type == "cuda"
        with torch.autocast(dev_type, enabled=amp):
            fmap1, fmap2 = self.fnet([image1, image2])
        corr_fn = self.corr_block(fmap1.float(), fmap2.float(), radius=4)
        with torch.autocast(dev_type, enabled=amp):
            net, inp = torch.split(self.cnet(image1), [128, 128], dim=1)
            net, inp = torch.tanh(net), torch.relu(inp)
            attention = self.att(inp)
        if (dev_type == "cuda" and torch.is_grad_enabled() and attention.requires_grad and self.args["num_heads"] == 1
                and os.environ.get("PCFA_GMA_ATTN_ACC", "1") != "0"):
            holder = {"shape": tuple(attention.shape)}
            attention = _AttnSource.apply(attention, holder)
            attention._pcfa_holder = holder
        N, _, H, W = image1.shape
        coords0 = coords_grid(N, H // 8, W // 8, image1.device)
        coords1 = coords0.clone()
        if flow_init is not None:
            coords1 = coords1 + flow_init
        predictions, flow_up = [], None
        from ..corr_block import CorrBlock as _OwnCorrBlock
        cl = (bool(getattr(self.update_block, "channels_last", False)) and dev_type == "cuda"
              and isinstance(corr_fn, _OwnCorrBlock))
        step_sources = None
        if cl:
            net = net.contiguous(memory_format=torch.channels_last)
            inp = inp.contiguous(memory_format=torch.channels_last)
            frozen = not any(p.requires_grad for p in self.update_block.gru.parameters())
            if frozen and torch.is_grad_enabled() and os.environ.get("PCFA_GRU_STEP", "1") != "0":
                from ..gru_ops import hoist_sources
                with torch.autocast(dev_type, enabled=amp):
                    step_sources = hoist_sources(self.update_block.gru.hoisted(inp))
        flow_cl = None
        for itr in range(iters):
            coords1 = coords1.detach()
            corr = corr_fn(coords1, channels_last=True) if cl else corr_fn(coords1)
            need_up = (not test_mode) or itr == iters - 1
            if cl:
                # zero-padded 8-channel flow (a tensor-core 7x7 convolution without cuDNN's channel padding) and ONE launch for
                # coords1 += delta, flow = coords1 - coords0 (conv_ops.flow_step)
                from ..conv_ops import flow_step, padded_flow
                if flow_cl is None:
                    flow_cl = padded_flow(coords1 - coords0, 8)
                with torch.autocast(dev_type, enabled=amp):
                    net, up_mask, delta_flow = self.update_block(net, inp, corr, flow_cl, attention, want_mask=need_up, raw_mask=True,
                                                                 step_sources=step_sources, last=itr == iters - 1)
                coords1, flow_cl = flow_step(coords1, coords0, delta_flow, 8)
            else:
                flow = coords1 - coords0
                with torch.autocast(dev_type, enabled=amp):
                    net, up_mask, delta_flow = self.update_block(net, inp, corr, flow, attention, want_mask=need_up, raw_mask=False,
                                                                 step_sources=step_sources, last=itr == iters - 1)
                coords1 = coords1 + delta_flow.float().contiguous()
            if need_up:
                if cl:                                 # fused kernel on the raw mask (csrc/upsample.cu), fp32 like the reference's
                    from ..upsample import convex_upsample as _fused_upsample      # fp32 softmax under autocast
                    flow_up = _fused_upsample(coords1 - coords0, up_mask, 0.25)
                else:
                    flow_up = convex_upsample(coords1 - coords0, up_mask)
                predictions.append(flow_up)
        if test_mode:
            return coords1 - coords0, flow_up
        return predictions
