"""Model adapter: padding, pre/post-processing and the ScaledInputModel wrapper.

Host-side mirror of helper_functions/ownutilities.py (InputPadder :21-62, preprocess_img :241-280,
postprocess_flow :283-299, compute_flow :302-343, model_takes_unit_input :347-360) and
helper_functions/own_models.py (ScaledInputModel :9-88), with two deliberate differences:
  * postprocess_flow does not move the flow to the host (the reference's `.cpu()` at :297 forces a
    device→host→device round trip in every closure evaluation);
  * weights come from `deterministic_state_` (no checkpoints are reachable offline); a checkpoint path
    can be passed instead.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import objective as J

NETWORKS = ("RAFT", "GMA", "PWCNet", "SpyNet", "FlowNet2")


class InputPadder:
    """Pads images such that dimensions are divisible by `divisor` (replicate padding)."""

    def __init__(self, dims, divisor=8, mode="sintel"):
        self.ht, self.wd = dims[-2:]
        pad_ht = (((self.ht // divisor) + 1) * divisor - self.ht) % divisor
        pad_wd = (((self.wd // divisor) + 1) * divisor - self.wd) % divisor
        if mode == "sintel":
            self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, pad_ht // 2, pad_ht - pad_ht // 2]
        else:
            self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, 0, pad_ht]

    @property
    def top_left(self):
        return self._pad[2], self._pad[0]

    def pad(self, *inputs):
        return [F.pad(x, self._pad, mode="replicate") for x in inputs]

    def get_dimensions(self):
        return self.ht, self.wd

    def unpad(self, x):
        ht, wd = x.shape[-2:]
        return x[..., self._pad[2]:ht - self._pad[3], self._pad[0]:wd - self._pad[1]]


def model_takes_unit_input(net: str) -> bool:
    return net in ("PWCNet", "SpyNet")


def preprocess_img(network, *images):
    if network in ("RAFT", "GMA"):
        padder = InputPadder(images[0].shape)
        return padder, padder.pad(*images)
    if network in ("PWCNet", "SpyNet"):
        images = [img / 255. for img in images]
        padder = InputPadder(images[0].shape, divisor=64)
        return padder, padder.pad(*images)
    if network[:7] == "FlowNet":
        if network[:8] != "FlowNet2":
            images = [img / 255. for img in images]
        padder = InputPadder(images[0].shape, divisor=64)
        return padder, padder.pad(*images)
    return None, images


def postprocess_flow(network, padder, *flows):
    """Remove the padding.  Stays on the device (the reference appends .cpu(), ownutilities.py:297)."""
    if padder is not None:
        return [padder.unpad(flow) for flow in flows]
    return flows


def compute_flow(model, network, x1, x2, test_mode=True, **kwargs):
    if network == "scaled_input_model":
        return model(x1, x2, test_mode=True, **kwargs)
    if network == "RAFT":
        return model(x1, x2, test_mode=test_mode, **kwargs)[1]
    if network == "GMA":
        return model(x1, x2, iters=6, test_mode=test_mode, **kwargs)[1]
    if network[:7] == "FlowNet":
        x = torch.stack((x1, x2), dim=-3)
        if network[:8] != "FlowNet2":
            mean = x.contiguous().view(x.size()[:2] + (-1,)).mean(dim=-1).view(x.size()[:2] + (1, 1, 1)).detach()
            x = x - mean
        return model(x)
    return model(x1, x2, **kwargs)


def build_network(net: str, device="cuda", seed: int = 0, weights: str | None = None, ops=None, gain: float = 1.0,
                  channels_last_encoders: bool = True, channels_last_update: bool = True):
    """Construct a flow network with this package's operators (or `ops`, used by the tests to inject
    the oracle), weights from `weights` (a checkpoint path) or deterministic synthetic values."""
    from .networks.weights import deterministic_state_
    if net == "RAFT":
        from .networks.raft import RAFT
        model = RAFT({"small": False, "mixed_precision": False}, corr_block=getattr(ops, "CorrBlock", None))
    elif net == "GMA":
        from .networks.gma import RAFTGMA
        model = RAFTGMA(corr_block=getattr(ops, "CorrBlock", None))
    elif net == "PWCNet":
        from .networks.pwcnet import PWCDCNet
        model = PWCDCNet(ops=ops)
    elif net == "FlowNet2":
        from .networks.flownet2 import FlowNet2
        model = FlowNet2(ops=ops)
    else:
        raise RuntimeWarning("The network %s is not a valid model option for import_and_load(network)." % net)
    if weights is not None:
        sd = torch.load(weights, map_location="cpu")
        model.load_state_dict(sd.get("state_dict", sd) if isinstance(sd, dict) else sd)
    else:
        deterministic_state_(model, seed, gain=gain)
    model = model.to(device).eval()
    for p in model.parameters():
        p.requires_grad = False
    if channels_last_encoders and torch.device(device).type == "cuda":
        for name in ("fnet", "cnet"):                       # RAFT / GMA encoders run NHWC end to end (networks/raft.py)
            enc = getattr(model, name, None)
            if enc is not None:
                enc.to(memory_format=torch.channels_last)
                enc.channels_last = True
        if net == "FlowNet2":                                   # plain conv stacks: NHWC weights make every activation NHWC
            model.to(memory_format=torch.channels_last)             # (closure 15.6 -> 11.7 ms, scripts/cl_experiment.py);
        if net == "PWCNet" and os.environ.get("PCFA_PWC_CL", "1") != "0":
            # round 1 measured plain channels-last weights as slower (5.9 -> 6.7 ms: ATen's channels-last cat, conversions
            # around the NCHW correlation / warp operators); with the vectorised cat kernel, one conversion per feature level
            # and the fused convolution epilogue the NHWC path wins (networks/pwcnet.py)
            model.to(memory_format=torch.channels_last)
            model.channels_last = True
        if net == "RAFT" and channels_last_update:             # NHWC update block: networks/raft.py, BasicUpdateBlock.forward
            model.update_block.to(memory_format=torch.channels_last)
            model.update_block.channels_last = True
        if net == "GMA" and channels_last_update:              # networks/gma.py: NHWC update block, attention on views
            model.update_block.to(memory_format=torch.channels_last)
            model.att.to(memory_format=torch.channels_last)
            model.update_block.channels_last = True
        if net == "GMA":
            from .networks.amp import install_frozen_half_weights
            install_frozen_half_weights(model)
    return model


class ScaledInputModel(nn.Module):
    """own_models.py:9-88 — takes [0,1] inputs (or the C&W variable w), optional deltas, and calls the
    wrapped network.  The elementwise pre-processing runs in the fused box kernel."""

    def __init__(self, net, make_unit_input=False, variable_change=False, model=None, ops=None, **kwargs):
        super().__init__()
        self.make_unit_input = make_unit_input
        self.var_change = variable_change
        self.model_name = net
        self.eps_box = kwargs.get("eps_box", 0.)
        self._ops = ops if ops is not None else J
        self.model_loaded = model if model is not None else build_network(
            net, device=kwargs.get("device", "cuda"), seed=kwargs.get("seed", 0), weights=kwargs.get("weights"))

    def forward(self, image1, image2, delta1=None, delta2=None, test_mode=True, *args, **kwargs):
        si = self._ops.scaled_input
        d2 = delta2 if delta2 is not None else delta1           # only delta1 given → added to both images
        x1 = si(image1, delta1, var_change=self.var_change, eps_box=self.eps_box, make_unit_input=self.make_unit_input)
        x2 = si(image2, d2, var_change=self.var_change, eps_box=self.eps_box, make_unit_input=self.make_unit_input)
        return compute_flow(self.model_loaded, self.model_name, x1, x2, test_mode=test_mode, *args, **kwargs)


def import_and_load(net="RAFT", make_unit_input=False, variable_change=False, device="cuda",
                    make_scaled_input_model=False, **kwargs):
    """ownutilities.py:64-169."""
    if make_unit_input or variable_change or make_scaled_input_model:
        return ScaledInputModel(net, make_unit_input=make_unit_input, variable_change=variable_change,
                                device=device, **kwargs)
    return build_network(net, device=device, seed=kwargs.get("seed", 0), weights=kwargs.get("weights"))
